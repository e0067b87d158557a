// Device-side scheduling of constraint groups into conflict-free phases.
//
// The reference sweeps every row sequentially (src/solver/sor_prox.rs:159-179).  Here rows are
// grouped (all rows of <= 4 contacts of one manifold / of one joint: they share their body pair,
// so one thread runs them in order) and groups are assigned to phases such that no two groups of
// a phase touch the same dynamic body:
//   * NB2_MODE_REFERENCE_ORDER: phase = level = 1 + max(level of the previous group, in the
//     reference's sweep order, on either body).  Executing levels in order reproduces the
//     sequential sweep exactly.
//   * NB2_MODE_COLOURED: phase = colour from a Jones-Plassmann greedy colouring of the group
//     conflict graph, computed on device with per-body colour bitmasks.
// Non-dynamic bodies (ground, kinematic) never create conflicts.
#include <stddef.h>
#include <stdlib.h>

#include "solver.cuh"

namespace nb2 {

static const int TPB = 256;
static inline unsigned int nblk(size_t n) { return (unsigned int)((n + TPB - 1) / TPB); }

// ------------------------------------------------------------------ exclusive scan
#define SCAN_TPB 512
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_TPB * SCAN_ITEMS)

__device__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* total) {
    __shared__ unsigned int warp_sums[SCAN_TPB / 32];
    unsigned int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned int s = lane < SCAN_TPB / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < SCAN_TPB / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    unsigned int base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[SCAN_TPB / 32 - 1];
    __syncthreads();
    return base + x - v;
}

__global__ void k_scan_tile_sums(const unsigned int* __restrict__ in, size_t n, unsigned int* tile_sums) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    unsigned int total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// single block: in-place exclusive scan of the tile sums (any count)
__global__ void k_scan_sums(unsigned int* tile_sums, unsigned int ntiles) {
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned int start = 0; start < ntiles; start += SCAN_TPB) {
        unsigned int i = start + threadIdx.x;
        unsigned int v = i < ntiles ? tile_sums[i] : 0;
        unsigned int total;
        unsigned int ex = block_exclusive_scan(v, &total);
        unsigned int c = carry;
        if (i < ntiles) tile_sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}
// writes out[0..n) = exclusive prefix and out[n] = total
__global__ void k_scan_apply(const unsigned int* __restrict__ in, size_t n, const unsigned int* __restrict__ tile_sums,
                             unsigned int* out) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned int v[SCAN_ITEMS];
    unsigned int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    unsigned int total;
    unsigned int ex = block_exclusive_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
        if (base + k + 1 == n) out[n] = ex;
    }
}
__global__ void k_zero_one(unsigned int* p) { *p = 0; }

int exclusive_scan_u32(Context* ctx, const unsigned int* in, unsigned int* out, size_t n) {
    if (n == 0) {
        k_zero_one<<<1, 1, 0, ctx->stream>>>(out);
        ctx->launches++;
        NB2_CUDA(ctx, cudaGetLastError());
        return NB2_OK;
    }
    unsigned int ntiles = (unsigned int)((n + SCAN_TILE - 1) / SCAN_TILE);
    NB2_TRY(ctx->scan_tmp.reserve(ctx, ntiles + 1));
    k_scan_tile_sums<<<ntiles, SCAN_TPB, 0, ctx->stream>>>(in, n, ctx->scan_tmp.p);
    k_scan_sums<<<1, SCAN_TPB, 0, ctx->stream>>>(ctx->scan_tmp.p, ntiles);
    k_scan_apply<<<ntiles, SCAN_TPB, 0, ctx->stream>>>(in, n, ctx->scan_tmp.p, out);
    ctx->launches += 3;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

// ------------------------------------------------------------------ input validation
// Bad records are neutralised on device (no host-side scan of the caller's arrays on the step
// path); the flag is reported by nb2_synchronize / nb2_get_stats / downloads.
__global__ void k_validate_manifolds(nb2_manifold* m, unsigned int nm, unsigned int nb, unsigned int nc,
                                     unsigned int* flags) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    nb2_manifold& mf = m[i];
    bool bad = mf.body1 < 0 || mf.body2 < 0 || (unsigned int)mf.body1 >= nb || (unsigned int)mf.body2 >= nb ||
               (unsigned long long)mf.first_contact + mf.num_contacts > nc;
    bool self = !bad && mf.body1 == mf.body2;
    if (bad || self) {
        mf.body1 = 0;
        mf.body2 = 0;
        mf.first_contact = 0;
        mf.num_contacts = 0;
        atomicOr(flags, bad ? 1u : 2u);
    }
}
__global__ void k_validate_joints(nb2_joint* j, unsigned int nj, unsigned int nb, unsigned int* flags) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nj) return;
    nb2_joint& jt = j[i];
    bool bad = jt.body1 < 0 || jt.body2 < 0 || (unsigned int)jt.body1 >= nb || (unsigned int)jt.body2 >= nb ||
               jt.type >= NB2_JOINT_TYPE_COUNT;
    bool self = !bad && jt.body1 == jt.body2;
    if (bad || self) {
        jt.body1 = 0;
        jt.body2 = 0;
        jt.type = NB2_JOINT_BALL;
        jt.broken = 1;
        atomicOr(flags, bad ? 1u : 2u);
    }
}
int launch_validate_inputs(Context* ctx) {
    NB2_TRY(ctx->flags.reserve(ctx, 4));
    if (ctx->n_manifolds) {
        k_validate_manifolds<<<nblk(ctx->n_manifolds), TPB, 0, ctx->stream>>>(ctx->manifolds.p, ctx->n_manifolds,
                                                                              ctx->n_bodies, ctx->n_contacts,
                                                                              ctx->flags.p);
        ctx->launches++;
    }
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}
int launch_validate_joints(Context* ctx) {
    NB2_TRY(ctx->flags.reserve(ctx, 4));
    if (ctx->n_joints) {
        k_validate_joints<<<nblk(ctx->n_joints), TPB, 0, ctx->stream>>>(ctx->joints.p, ctx->n_joints, ctx->n_bodies,
                                                                        ctx->flags.p);
        ctx->launches++;
    }
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

// ------------------------------------------------------------------ items
// min_chunks = 1 for the device producer's manifolds (one per persistent pair, <= 4 contacts each): a pair
// that loses its contacts keeps an (unscheduled) chunk, so the item numbering -- what the schedule cache and
// the incremental recolouring compare from step to step -- does not shift.
__global__ void k_chunk_counts(const nb2_manifold* __restrict__ m, unsigned int nm, unsigned int* counts,
                               unsigned int min_chunks) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nm) counts[i] = max((m[i].num_contacts + NB2_CHUNK - 1) / NB2_CHUNK, min_chunks);
}
__global__ void k_fill_chunks(const nb2_manifold* __restrict__ m, unsigned int nm,
                              const unsigned int* __restrict__ chunk_base, unsigned int* chunk_manifold,
                              unsigned int* c_manifold, unsigned int maxc, unsigned int* flags) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    unsigned int cb = chunk_base[i], ce = chunk_base[i + 1];
    // manifolds whose contact ranges overlap can claim more chunks than the n_manifolds + n_contacts / 4
    // the buffers were sized for: the excess is dropped and reported as a bad record
    if (ce > maxc) {
        atomicOr(flags, 1u);
        ce = maxc;
    }
    for (unsigned int k = cb; k < ce; ++k) chunk_manifold[k] = i;
    unsigned int f = m[i].first_contact, nc = m[i].num_contacts;
    for (unsigned int k = 0; k < nc; ++k) c_manifold[f + k] = i;
}

__device__ __forceinline__ int dyn_or_neg(const int* status, int b) {
    return status[b] == NB2_BODY_DYNAMIC ? b : -1;
}
__device__ __forceinline__ int joint_max_rows(unsigned int type) {
    // num_velocity_constraints() of each *_constraint.rs
    const int rows[NB2_JOINT_TYPE_COUNT] = {3, 5, 7, 4, 3, 4, 4, 4, 6, 3};
    return rows[type];
}
__device__ __forceinline__ int joint_num_position(const nb2_joint& j) {
    switch (j.type) {
        case NB2_JOINT_BALL:
        case NB2_JOINT_CARTESIAN: return 1;
        case NB2_JOINT_PRISMATIC:
            return (j.flags & (NB2_JOINT_FLAG_MIN_OFFSET | NB2_JOINT_FLAG_MAX_OFFSET)) ? 3 : 2;
        default: return 2;
    }
}

// One thread per potential item.  Velocity schedule layout:
//   [0, nJ) joints | [nJ, nJ+maxc) contact chunks (coloured: all rows; reference: friction rows)
//   | [nJ+maxc, nJ+2maxc) reference order only: normal rows of the chunks
// Position schedule (reference order only): [0,nJ) joints | [nJ, nJ+maxc) chunks.
// rpc = velocity rows per contact: 3 (normal + two friction rows) or 1 (frictionless SignoriniModel)
__global__ void k_build_items(int mode, int compact, int rpc, int position, unsigned int nJ, unsigned int maxc,
                              const nb2_joint* __restrict__ joints, const nb2_manifold* __restrict__ manifolds,
                              unsigned int* chunk_base, unsigned int nM, unsigned int* chunk_manifold,
                              unsigned int* c_manifold, int producer, const int* __restrict__ status,
                              int* it_a, int* it_b, int* it_nrows, int* it_type, int* it_src,
                              unsigned long long* it_key, int* it_b1, int* it_b2, size_t n_items) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    int a = -1, b = -1, nrows = 0, type = NB2_ITEM_INVALID, src = 0, b1 = 0, b2 = 0;
    unsigned long long key = 0;
    if (i < nJ) {
        const nb2_joint& j = joints[i];
        b1 = j.body1;
        b2 = j.body2;
        a = dyn_or_neg(status, j.body1);
        b = dyn_or_neg(status, j.body2);
        // active_joints filter of mechanical_world.rs:274-279 (+ is_active, joint_constraint.rs:219-228)
        if (!j.broken && (a >= 0 || b >= 0)) {
            type = NB2_ITEM_JOINT;
            src = (int)i;
            nrows = position ? joint_num_position(j) : joint_max_rows(j.type);
            unsigned long long bucket = position ? 0ull : ((a < 0 || b < 0) ? 1ull : 0ull);
            key = (bucket << 40) | (unsigned long long)i;
        }
    } else {
        size_t k = i - nJ;
        int second = 0;
        if (k >= maxc) {
            k -= maxc;
            second = 1;
        }
        // The device producer's manifolds own exactly one chunk and four contact slots each (narrowphase.cu):
        // chunk k IS manifold k, and the chunk bookkeeping (a memset, k_chunk_counts, three scan kernels and
        // k_fill_chunks for uploaded manifolds) is written right here.
        unsigned int total = producer ? nM : chunk_base[nM];
        if (k < total) {
            unsigned int m = producer ? (unsigned int)k : chunk_manifold[k];
            const nb2_manifold& mf = manifolds[m];
            if (producer && !second && !position) {
                chunk_base[m] = m;
                if (m == 0) chunk_base[nM] = nM;
                chunk_manifold[m] = m;
                const unsigned int nc = mf.num_contacts;
                *reinterpret_cast<uint4*>(c_manifold + 4u * m) =
                    make_uint4(nc > 0 ? m : 0xFFFFFFFFu, nc > 1 ? m : 0xFFFFFFFFu, nc > 2 ? m : 0xFFFFFFFFu, nc > 3 ? m : 0xFFFFFFFFu);
            }
            b1 = mf.body1;
            b2 = mf.body2;
            a = dyn_or_neg(status, mf.body1);
            b = dyn_or_neg(status, mf.body2);
            const int local = producer ? 0 : (int)(k - chunk_base[m]);
            const int ncc = min(NB2_CHUNK, (int)mf.num_contacts - NB2_CHUNK * local);
            if ((a >= 0 || b >= 0) && ncc > 0) {
                src = (int)k;
                bool ground = (a < 0 || b < 0);
                if (position) {
                    type = NB2_ITEM_CONTACTS;
                    nrows = ncc;
                    key = (1ull << 40) | k;
                } else if (mode == NB2_MODE_COLOURED) {
                    // compact contact groups (c_geo planes) own no generic row slots; their
                    // contact count travels in bits 4..7
                    type = NB2_ITEM_CONTACTS;
                    nrows = compact ? (ncc << 4) : rpc * ncc;
                } else if (!second) {
                    if (rpc == 3) {  // the frictionless model has no friction bucket
                        type = NB2_ITEM_FRICTION;
                        nrows = 2 * ncc;
                        key = ((ground ? 3ull : 2ull) << 40) | k;
                    }
                } else {
                    type = NB2_ITEM_NORMAL;
                    nrows = ncc;
                    key = ((ground ? 5ull : 4ull) << 40) | k;
                }
            }
        }
    }
    it_a[i] = a;
    it_b[i] = b;
    it_nrows[i] = nrows;
    it_type[i] = type;
    it_src[i] = src;
    it_key[i] = key;
    it_b1[i] = b1;
    it_b2[i] = b2;
}

static int reserve_sched(Context* ctx, Sched* s, size_t n_items, size_t max_phases) {
    s->n_items = n_items;
    s->max_phases = max_phases;
    NB2_TRY(s->it_a.reserve(ctx, n_items));
    NB2_TRY(s->it_b.reserve(ctx, n_items));
    NB2_TRY(s->it_nrows.reserve(ctx, n_items));
    NB2_TRY(s->it_type.reserve(ctx, n_items));
    NB2_TRY(s->it_src.reserve(ctx, n_items));
    NB2_TRY(s->it_b1.reserve(ctx, n_items));
    NB2_TRY(s->it_b2.reserve(ctx, n_items));
    NB2_TRY(s->it_key.reserve(ctx, n_items));
    NB2_TRY(s->it_phase.reserve(ctx, n_items));
    NB2_TRY(s->it_slot.reserve(ctx, n_items));
    NB2_TRY(s->ph_count.reserve(ctx, max_phases + 1));
    NB2_TRY(s->ph_R.reserve(ctx, max_phases + 1));
    NB2_TRY(s->ph_bcnt.reserve(ctx, 4 * (max_phases + 1)));
    NB2_TRY(s->ph_gbase.reserve(ctx, max_phases + 1));
    NB2_TRY(s->ph_rbase.reserve(ctx, max_phases + 1));
    NB2_TRY(s->g_info.reserve(ctx, n_items));
    {
        const SchedHeader* before = s->hdr.p;
        NB2_TRY(s->hdr.reserve(ctx, 1));
        if (s->hdr.p != before) NB2_CUDA(ctx, cudaMemsetAsync(s->hdr.p, 0, s->hdr.cap * sizeof(SchedHeader), ctx->stream));  // padding words too
    }
    return NB2_OK;
}

int launch_build_items(Context* ctx, int mode) {
    const unsigned int nM = ctx->n_manifolds, nJ = ctx->n_joints;
    const size_t maxc = ctx->max_chunks;
    // chunk bookkeeping
    NB2_TRY(ctx->chunk_base.reserve(ctx, nM + 2));
    NB2_TRY(ctx->chunk_manifold.reserve(ctx, maxc + 1));
    NB2_TRY(ctx->c_manifold.reserve(ctx, ctx->n_contacts + 1));
    NB2_TRY(ctx->deg.reserve(ctx, (size_t)nM + 1));  // reused as scan input
    // the producer's layout (one chunk, four contact slots per manifold) is booked by k_build_items itself
    const bool producer = ctx->manifolds_from_producer && nM > 0 && ctx->n_contacts == 4u * nM && maxc == (size_t)nM;
    // contacts not owned by a (valid) manifold stay unmapped and are skipped by assembly
    if (!producer)
        NB2_CUDA(ctx, cudaMemsetAsync(ctx->c_manifold.p, 0xFF, ((size_t)ctx->n_contacts + 1) * sizeof(unsigned int),
                                      ctx->stream));
    if (nM && !producer) {
        k_chunk_counts<<<nblk(nM), TPB, 0, ctx->stream>>>(ctx->manifolds.p, nM, ctx->deg.p, ctx->manifolds_from_producer ? 1u : 0u);
        ctx->launches++;
    }
    if (!producer) NB2_TRY(exclusive_scan_u32(ctx, ctx->deg.p, ctx->chunk_base.p, nM));
    if (nM && !producer) {
        k_fill_chunks<<<nblk(nM), TPB, 0, ctx->stream>>>(ctx->manifolds.p, nM, ctx->chunk_base.p,
                                                         ctx->chunk_manifold.p, ctx->c_manifold.p,
                                                         (unsigned int)maxc, ctx->flags.p);
        ctx->launches++;
    }
    const bool ref = mode == NB2_MODE_REFERENCE_ORDER;
    size_t n_items = nJ + (ref ? 2 : 1) * maxc;
    NB2_TRY(reserve_sched(ctx, &ctx->vs, n_items, ref ? n_items : NB2_MAX_COLOURS));
    if (n_items) {
        k_build_items<<<nblk(n_items), TPB, 0, ctx->stream>>>(
            mode, ctx->step_layout, ctx->contact_model == 1 ? 1 : 3, 0, nJ, (unsigned int)maxc, ctx->joints.p, ctx->manifolds.p, ctx->chunk_base.p, nM,
            ctx->chunk_manifold.p, ctx->c_manifold.p, producer ? 1 : 0, ctx->b_status.p, ctx->vs.it_a.p, ctx->vs.it_b.p, ctx->vs.it_nrows.p,
            ctx->vs.it_type.p, ctx->vs.it_src.p, ctx->vs.it_key.p, ctx->vs.it_b1.p, ctx->vs.it_b2.p, n_items);
        ctx->launches++;
    }
    if (ref) {
        size_t np = nJ + maxc;
        NB2_TRY(reserve_sched(ctx, &ctx->ps, np, np));
        if (np) {
            k_build_items<<<nblk(np), TPB, 0, ctx->stream>>>(
                mode, ctx->step_layout, ctx->contact_model == 1 ? 1 : 3, 1, nJ, (unsigned int)maxc, ctx->joints.p, ctx->manifolds.p, ctx->chunk_base.p, nM,
                ctx->chunk_manifold.p, ctx->c_manifold.p, producer ? 1 : 0, ctx->b_status.p, ctx->ps.it_a.p, ctx->ps.it_b.p, ctx->ps.it_nrows.p,
                ctx->ps.it_type.p, ctx->ps.it_src.p, ctx->ps.it_key.p, ctx->ps.it_b1.p, ctx->ps.it_b2.p, np);
            ctx->launches++;
        }
    }
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

// ------------------------------------------------------------------ reference order: levels
__global__ void k_degree(const int* __restrict__ it_a, const int* __restrict__ it_b, const int* __restrict__ it_type,
                         size_t n, unsigned int* deg) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || it_type[i] == NB2_ITEM_INVALID) return;
    int a = it_a[i], b = it_b[i];
    if (a >= 0) atomicAdd(&deg[a], 1u);
    if (b >= 0 && b != a) atomicAdd(&deg[b], 1u);
}
__global__ void k_fill_adj(const int* __restrict__ it_a, const int* __restrict__ it_b, const int* __restrict__ it_type,
                           size_t n, const unsigned int* __restrict__ adj_off, unsigned int* cursor, int* adj) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || it_type[i] == NB2_ITEM_INVALID) return;
    int a = it_a[i], b = it_b[i];
    if (a >= 0) adj[adj_off[a] + atomicAdd(&cursor[a], 1u)] = (int)i;
    if (b >= 0 && b != a) adj[adj_off[b] + atomicAdd(&cursor[b], 1u)] = (int)i;
}
// one thread per body: order its incident groups by sweep key, then link each group to its
// predecessor on this body
__global__ void k_sort_adj_link(unsigned int nb, const unsigned int* __restrict__ adj_off, int* adj,
                                const unsigned long long* __restrict__ key, const int* __restrict__ it_a,
                                int* pred_a, int* pred_b) {
    unsigned int body = blockIdx.x * blockDim.x + threadIdx.x;
    if (body >= nb) return;
    unsigned int s = adj_off[body], e = adj_off[body + 1];
    for (unsigned int i = s + 1; i < e; ++i) {  // insertion sort (segments are short)
        int v = adj[i];
        unsigned long long kv = key[v];
        unsigned int j = i;
        while (j > s && key[adj[j - 1]] > kv) {
            adj[j] = adj[j - 1];
            --j;
        }
        adj[j] = v;
    }
    int prev = -1;
    for (unsigned int i = s; i < e; ++i) {
        int v = adj[i];
        if (it_a[v] == (int)body)
            pred_a[v] = prev;
        else
            pred_b[v] = prev;
        prev = v;
    }
}
__global__ void k_fill_int(int* p, size_t n, int v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// cooperative: iterate level[i] = 1 + max(level[pred]) to the fixed point
__global__ void __launch_bounds__(TPB) k_levelise(size_t n, const int* __restrict__ it_type,
                                                  const int* __restrict__ pred_a, const int* __restrict__ pred_b,
                                                  int* level, unsigned int* flags /*3*/, unsigned int* barrier) {
    GridBarrier gb;
    gb.init(barrier);
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (unsigned int round = 0;; ++round) {
        unsigned int changed = 0;
        for (size_t i = tid; i < n; i += stride) {
            if (it_type[i] == NB2_ITEM_INVALID) continue;
            int pa = pred_a[i], pb = pred_b[i];
            int la = pa >= 0 ? __ldcg(&level[pa]) : -1;
            int lb = pb >= 0 ? __ldcg(&level[pb]) : -1;
            int nl = max(la, lb) + 1;
            if (nl > __ldcg(&level[i])) {
                __stcg(&level[i], nl);
                changed = 1;
            }
        }
        if (__syncthreads_or(changed) && threadIdx.x == 0) atomicOr(&flags[round % 3], 1u);
        gb.sync();
        unsigned int any = *((volatile unsigned int*)&flags[round % 3]);
        if (tid == 0) flags[(round + 2) % 3] = 0;
        if (!any) break;
    }
}

// ------------------------------------------------------------------ coloured: Jones-Plassmann
__device__ __forceinline__ unsigned int hash_u32(unsigned int x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
// Schedule-cache verdicts (the device flag `changed`): 0 keep everything, 1 (or ~0) colour from scratch, 2 the
// graph is unchanged but a refinement pass is still due, 3 a few groups changed: edit the colouring in place.
#define NB2_SCHED_REFINE 2u
#define NB2_SCHED_INCREMENTAL 3u
// Compares this step's groups with the previous step's (when comparable): counts the groups whose conflict
// signature -- (dynamic body pair, row count, type, body pair) -- differs.
__global__ void k_compare_snapshot(size_t n, const int* __restrict__ it_a, const int* __restrict__ it_b,
                                   const int* __restrict__ it_nrows, const int* __restrict__ it_type,
                                   const int* __restrict__ it_b1, const int* __restrict__ it_b2,
                                   const int* __restrict__ prev_a, const int* __restrict__ prev_b,
                                   const int* __restrict__ prev_nt, const int* __restrict__ prev_b1,
                                   const int* __restrict__ prev_b2, unsigned int* n_changed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool diff = false;
    if (i < n) {
        const int nt = (it_nrows[i] & 0xFF) | (it_type[i] << 8);
        diff = prev_a[i] != it_a[i] || prev_b[i] != it_b[i] || prev_nt[i] != nt || prev_b1[i] != it_b1[i] || prev_b2[i] != it_b2[i];
    }
    const unsigned int v = __ballot_sync(0xffffffffu, diff);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(n_changed, (unsigned int)__popc(v));
}
// Records this step's groups for the next comparison.  On an incremental step (*changed == NB2_SCHED_INCREMENTAL)
// it first edits the colouring in place: a group that vanished or changed its bodies gives its colour back
// (its bit is cleared in both bodies' colour masks), a group that appeared or changed its bodies is marked
// uncoloured for the Jones-Plassmann rounds of k_colour; a group that only changed its row count keeps its
// colour (conflicts depend on the bodies alone).
__global__ void k_apply_snapshot(size_t n, const int* __restrict__ it_a, const int* __restrict__ it_b,
                                 const int* __restrict__ it_nrows, const int* __restrict__ it_type,
                                 const int* __restrict__ it_b1, const int* __restrict__ it_b2, int* prev_a, int* prev_b,
                                 int* prev_nt, int* prev_b1, int* prev_b2, const unsigned int* __restrict__ changed,
                                 int* phase, unsigned long long* cmask, const unsigned int* __restrict__ ph_count_prev,
                                 const SchedHeader* __restrict__ hdr) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = it_a[i], b = it_b[i], nt = (it_nrows[i] & 0xFF) | (it_type[i] << 8), b1 = it_b1[i], b2 = it_b2[i];
    if (*changed == NB2_SCHED_INCREMENTAL) {
        const int pa = prev_a[i], pb = prev_b[i];
        const bool was = (prev_nt[i] >> 8) != NB2_ITEM_INVALID, now = it_type[i] != NB2_ITEM_INVALID;
        bool same = pa == a && pb == b;
        // A colour that holds almost nothing (groups that found no room when they appeared) still costs the
        // solve kernels a barrier and a latency chain per sweep: its groups try again every step -- room opens
        // up as other groups vanish.  (Read before this step's layout kernels clear the previous counts.)
        if (was && now && same) {
            const int c = phase[i];
            const unsigned int np_prev = hdr->n_phases, ng_prev = hdr->n_groups;
            if (c >= 0 && (unsigned int)c < np_prev && np_prev > 1u && ph_count_prev[c] * 16u * np_prev < ng_prev) same = false;
        }
        if (was && !(now && same)) {
            const int c = phase[i];
            if (c >= 0 && c < NB2_MAX_COLOURS) {
                const unsigned long long m = ~(1ull << (c & 63));
                if (pa >= 0) atomicAnd(&cmask[(size_t)pa * NB2_MASK_WORDS + (c >> 6)], m);
                if (pb >= 0) atomicAnd(&cmask[(size_t)pb * NB2_MASK_WORDS + (c >> 6)], m);
            }
        }
        if (now && !(was && same)) phase[i] = -1;
    }
    prev_a[i] = a;
    prev_b[i] = b;
    prev_nt[i] = nt;
    prev_b1[i] = b1;
    prev_b2[i] = b2;
}

// Schedule-cache verdict of the step: *changed = 0 keep everything, 1 (or ~0) colour from scratch, 2 = the
// graph is unchanged but the colouring still has refinement passes to run (one iterated-greedy pass per
// step, so a scene pays for them only while it stays put).
// ... 3 = a few groups changed: the colouring is edited in place (k_apply_snapshot + the Jones-Plassmann
// rounds over the uncoloured groups only; no iterated-greedy pass, no balancing).  Every NB2_INC_MAX-th
// incremental step in a row, and any step that changes more than 1/8 of the groups, colours from scratch.
#define NB2_INC_MAX 64u
__global__ void k_refine_decide(unsigned int* changed, SchedHeader* hdr, const unsigned int* __restrict__ n_changed,
                                unsigned int n_items, int comparable, int incremental) {
    const unsigned int nc = *n_changed;
    if (!comparable) {
        *changed = 1u;
    } else if (nc == 0u) {
        *changed = 0u;
    } else if (incremental && hdr->n_phases > 0u && hdr->overflow == 0u && nc <= max(256u, n_items / 8u) &&
               hdr->pad[0] < NB2_INC_MAX) {
        *changed = NB2_SCHED_INCREMENTAL;
    } else {
        *changed = 1u;
    }
    if (*changed == 1u) {
        hdr->refine_left = NB2_IG_REFINE;
        hdr->pad[0] = 0u;
    } else if (*changed == NB2_SCHED_INCREMENTAL) {
        hdr->refine_left = 0u;  // the pre-balance colouring the refinement passes work on is stale now
        hdr->pad[0] += 1u;
    } else if (hdr->refine_left > 0u) {
        *changed = NB2_SCHED_REFINE;
        hdr->refine_left -= 1u;
    }
}
// One launch for all the conditional resets of a scheduling pass (seven launches of 2.5 us each before, on
// every step, cached or not).  kind 0: zero unless the schedule is kept; 1: also kept on an incremental step
// (the colour masks); 2: fill with -1 unless the colours are carried over (refinement / incremental).
struct CondRegions {
    unsigned int* p[8];
    unsigned long long end[8];  // running end offsets, in words
    int kind[8];
    int n;
};
static void cond_add(CondRegions& cr, void* p, size_t nwords, int kind) {
    cr.p[cr.n] = (unsigned int*)p;
    cr.end[cr.n] = (cr.n ? cr.end[cr.n - 1] : 0ull) + nwords;
    cr.kind[cr.n] = kind;
    ++cr.n;
}
__global__ void k_cond_reset(const unsigned int* __restrict__ changed, CondRegions cr) {
    const unsigned int ch = *changed;
    if (ch == 0u) return;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long begin = 0ull;
    for (int k = 0; k < cr.n; ++k) {
        if (i < cr.end[k]) {
            if (cr.kind[k] == 1 && ch == NB2_SCHED_INCREMENTAL) return;
            if (cr.kind[k] == 2) {
                if (ch == NB2_SCHED_REFINE || ch == NB2_SCHED_INCREMENTAL) return;
                cr.p[k][i - begin] = 0xFFFFFFFFu;
                return;
            }
            cr.p[k][i - begin] = 0u;
            return;
        }
        begin = cr.end[k];
    }
}
static void cond_launch(Context* ctx, const unsigned int* changed, const CondRegions& cr) {
    const size_t total = (size_t)cr.end[cr.n - 1];
    k_cond_reset<<<(unsigned int)((total + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(changed, cr);
    ctx->launches++;
}


// ------------------------------------------------------------------ Kempe chains
// First-fit + iterated greedy leave the pile with max body degree 6 at 8 colours whose last class holds a few
// hundred groups (3718 after the fresh colouring, 280 after the refinement passes: tools/colour_classes.py).
// Every colour costs each sweep of both solve kernels a grid barrier and a latency chain, so that class is
// emptied by Kempe-chain interchanges (the constructive step of Vizing's / Konig's edge-colouring proofs): for a
// group e = (u, v) of the last class pick a colour `a` free at u and `b` free at v; the path of groups coloured
// a, b, a, ... that starts at v ends somewhere else than u (always, in a bipartite contact graph), swapping a and b
// along it frees `a` at v as well, and e takes it.  Parallel version, Jones-Plassmann style: every candidate bids
// for the bodies of its path, the winners of a round (disjoint paths) swap, the losers try again.
#define NB2_KEMPE_K 16          // colours the per-body table holds (the stage is skipped beyond)
#define NB2_KEMPE_MAX 8192u     // largest class the stage takes on
#define NB2_KEMPE_ROUNDS 10
#define NB2_KEMPE_LEN 96        // longest path followed
struct KempePlan {
    int kind;  // 0 nothing to do this round, 1 a colour free at both ends, 2 interchange along a path
    int u, v, a, b, n;
};
__device__ __forceinline__ unsigned int kempe_free(const unsigned long long* cmask, int body, unsigned int L) {
    const unsigned long long used = __ldcg(&cmask[(size_t)body * NB2_MASK_WORDS]);
    return (unsigned int)(~used) & ((1u << L) - 1u);
}
__device__ __forceinline__ int kempe_nth_bit(unsigned int m, unsigned int k) {  // k-th (mod popcount) set bit of m != 0
    k %= (unsigned int)__popc(m);
    for (unsigned int j = 0; j < k; ++j) m &= m - 1u;
    return __ffs((int)m) - 1;
}
// The plan of candidate e for round r.  Pure function of the tables as they stand after the last grid barrier, so
// the bidding pass and the applying pass of a round compute the same plan.  Visits the path's bodies through `visit`.
template <typename Visit>
__device__ KempePlan kempe_plan(int e, unsigned int r, unsigned int L, const int* __restrict__ it_a,
                                const int* __restrict__ it_b, const unsigned long long* cmask, const int* col_edge,
                                Visit visit) {
    KempePlan P;
    P.kind = 0;
    P.u = it_a[e];
    P.v = it_b[e];
    P.a = P.b = -1;
    P.n = 0;
    if (P.u < 0 || P.v < 0) {  // one dynamic body: any colour free there will do
        const int x = P.u < 0 ? P.v : P.u;
        if (x < 0) return P;
        const unsigned int f = kempe_free(cmask, x, L);
        if (!f) return P;
        P.kind = 1;
        P.a = __ffs((int)f) - 1;
        visit(x);
        return P;
    }
    const unsigned int fu = kempe_free(cmask, P.u, L), fv = kempe_free(cmask, P.v, L);
    if (fu & fv) {
        P.kind = 1;
        P.a = __ffs((int)(fu & fv)) - 1;
        visit(P.u);
        visit(P.v);
        return P;
    }
    if (!fu || !fv) return P;
    const unsigned int h = hash_u32((unsigned int)e * 0x9E3779B9u + r);
    P.a = kempe_nth_bit(fu, h + r);
    P.b = kempe_nth_bit(fv, (h >> 8) + r / 2u);
    int x = P.v, want = P.a, n = 0;
    for (;;) {
        const int f = __ldcg(&col_edge[(size_t)x * NB2_KEMPE_K + want]);
        if (f < 0) break;
        if (n >= NB2_KEMPE_LEN) return P;
        const int fa = it_a[f], fb = it_b[f];
        const int y = fa == x ? fb : fa;
        ++n;
        if (y < 0) break;          // a static side ends the path: nothing to keep proper there
        if (y == P.u) return P;    // odd cycle through e: this pair of colours does not work
        visit(y);
        x = y;
        want = want == P.a ? P.b : P.a;
    }
    if (n == 0) return P;  // cannot happen with exact tables
    visit(P.u);
    visit(P.v);
    P.kind = 2;
    P.n = n;
    return P;
}
__global__ void __launch_bounds__(TPB) k_colour(size_t n, const int* __restrict__ it_a, const int* __restrict__ it_b,
                                                const int* __restrict__ it_type, int* phase,
                                                unsigned long long* cmask, unsigned long long* best,
                                                unsigned int* flags /*3*/, SchedHeader* hdr, unsigned int* barrier,
                                                const unsigned int* __restrict__ changed, unsigned int* bal,
                                                int* tmp_phase, size_t nb, int ig_passes, int* raw_phase,
                                                int* col_edge) {
    if (*changed == 0u) return;  // cached colouring still valid (uniform over the grid: no barrier was touched)
    GridBarrier gb;
    gb.init(barrier);
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (unsigned int round = 1;; ++round) {
        // A: every uncoloured group bids on its bodies
        unsigned int pending = 0;
        for (size_t i = tid; i < n; i += stride) {
            if (it_type[i] == NB2_ITEM_INVALID || phase[i] >= 0) continue;
            pending = 1;
            unsigned long long prio = ((unsigned long long)round << 52) |
                                      ((unsigned long long)(hash_u32((unsigned int)i) & 0xFFFFFu) << 32) |
                                      (unsigned long long)(unsigned int)i;
            int a = it_a[i], b = it_b[i];
            if (a >= 0) atomicMax(&best[a], prio);
            if (b >= 0) atomicMax(&best[b], prio);
        }
        if (__syncthreads_or(pending) && threadIdx.x == 0) atomicOr(&flags[round % 3], 1u);
        gb.sync();
        unsigned int any = *((volatile unsigned int*)&flags[round % 3]);
        if (tid == 0) flags[(round + 2) % 3] = 0;
        if (!any) break;
        // B: the winner on both of its bodies takes the lowest free colour
        for (size_t i = tid; i < n; i += stride) {
            if (it_type[i] == NB2_ITEM_INVALID || phase[i] >= 0) continue;
            unsigned long long prio = ((unsigned long long)round << 52) |
                                      ((unsigned long long)(hash_u32((unsigned int)i) & 0xFFFFFu) << 32) |
                                      (unsigned long long)(unsigned int)i;
            int a = it_a[i], b = it_b[i];
            if (a >= 0 && __ldcg(&best[a]) != prio) continue;
            if (b >= 0 && __ldcg(&best[b]) != prio) continue;
            int colour = -1;
#pragma unroll
            for (int w = 0; w < NB2_MASK_WORDS; ++w) {
                unsigned long long ua = 0, ub = 0;
                if (a >= 0) ua = __ldcg(&cmask[(size_t)a * NB2_MASK_WORDS + w]);
                if (b >= 0) ub = __ldcg(&cmask[(size_t)b * NB2_MASK_WORDS + w]);
                const unsigned long long used = ua | ub;
                if (colour < 0 && used != ~0ull) {
                    int bit = __ffsll((long long)~used) - 1;
                    colour = w * 64 + bit;
                    unsigned long long m = 1ull << bit;
                    if (a >= 0) __stcg(&cmask[(size_t)a * NB2_MASK_WORDS + w], ua | m);
                    if (b >= 0) __stcg(&cmask[(size_t)b * NB2_MASK_WORDS + w], ub | m);
                }
            }
            if (colour < 0) {
                colour = NB2_MAX_COLOURS - 1;
                atomicOr(&hdr->overflow, 1u);
            }
            phase[i] = colour;
        }
        gb.sync();
    }
    if (*changed == NB2_SCHED_INCREMENTAL) {  // edited in place: the other groups keep their (refined, balanced) colours
        for (size_t i = tid; i < n; i += stride)
            if (it_type[i] != NB2_ITEM_INVALID) raw_phase[i] = phase[i];
        return;
    }
    // ---- iterated greedy (Culberson).  Groups of one colour share no body, so a whole colour class can be
    // recoloured at once with plain first-fit against the classes already redone; visiting the classes in
    // any order never needs more colours than before and often fewer.  Orders tried in turn: reverse,
    // largest class first, reverse.  Colours cost the solve kernels one barrier + one latency chain each.
    // fresh colouring: NB2_IG_PASSES passes now; refinement step: one pass, its order picked by how many are left
    const bool refine = *changed == NB2_SCHED_REFINE;
    const int ig_first = refine ? (int)(NB2_IG_REFINE - hdr->refine_left) : 0;
    const int ig_last = refine ? ig_first + 1 : ig_passes;
    // the passes work on the colouring as it was BEFORE balancing (kept in raw_phase): balanced classes are
    // all equally full and none of them can be absorbed any more
    if (refine) {
        for (size_t i = tid; i < n; i += stride)
            if (it_type[i] != NB2_ITEM_INVALID) phase[i] = raw_phase[i];
        gb.sync();
    }
    for (int pass = ig_first; pass < ig_last; ++pass) {
        __shared__ unsigned int s_order[NB2_MAX_COLOURS];
        __shared__ unsigned int s_ig_cnt[NB2_MAX_COLOURS];
        __shared__ unsigned int s_ig_C;
        for (size_t i = tid; i < NB2_MAX_COLOURS; i += stride) bal[i] = 0u;
        for (size_t w = tid; w < nb * (size_t)NB2_MASK_WORDS; w += stride) cmask[w] = 0ull;
        gb.sync();
        for (size_t i = tid; i < n; i += stride)
            if (it_type[i] != NB2_ITEM_INVALID) atomicAdd(&bal[min((unsigned int)phase[i], (unsigned int)NB2_MAX_COLOURS - 1u)], 1u);
        gb.sync();
        for (unsigned int c = threadIdx.x; c < NB2_MAX_COLOURS; c += blockDim.x) s_ig_cnt[c] = __ldcg(&bal[c]);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int C = 0;
            for (unsigned int c = 0; c < NB2_MAX_COLOURS; ++c)
                if (s_ig_cnt[c]) C = c + 1;
            s_ig_C = C;
            if (pass % 3 == 1) {  // largest class first (selection sort on the counts; ties by index)
                for (unsigned int k = 0; k < C; ++k) {
                    unsigned int best_c = 0, best_n = 0;
                    bool found = false;
                    for (unsigned int c = 0; c < C; ++c) {
                        if (s_ig_cnt[c] == 0xFFFFFFFFu) continue;
                        if (!found || s_ig_cnt[c] > best_n) {
                            best_c = c;
                            best_n = s_ig_cnt[c];
                            found = true;
                        }
                    }
                    s_order[k] = best_c;
                    s_ig_cnt[best_c] = 0xFFFFFFFFu;
                }
            } else {
                for (unsigned int k = 0; k < C; ++k) s_order[k] = C - 1 - k;
            }
        }
        __syncthreads();
        const unsigned int C = s_ig_C;
        for (unsigned int k = 0; k < C; ++k) {
            const int c = (int)s_order[k];
            for (size_t i = tid; i < n; i += stride) {
                if (it_type[i] == NB2_ITEM_INVALID || phase[i] != c) continue;
                const int a = it_a[i], b = it_b[i];
                int colour = -1;
#pragma unroll
                for (int w = 0; w < NB2_MASK_WORDS; ++w) {
                    unsigned long long ua = 0, ub = 0;
                    if (a >= 0) ua = __ldcg(&cmask[(size_t)a * NB2_MASK_WORDS + w]);
                    if (b >= 0) ub = __ldcg(&cmask[(size_t)b * NB2_MASK_WORDS + w]);
                    const unsigned long long used = ua | ub;
                    if (colour < 0 && used != ~0ull) {
                        const int bit = __ffsll((long long)~used) - 1;
                        colour = w * 64 + bit;
                        const unsigned long long m = 1ull << bit;
                        if (a >= 0) __stcg(&cmask[(size_t)a * NB2_MASK_WORDS + w], ua | m);
                        if (b >= 0) __stcg(&cmask[(size_t)b * NB2_MASK_WORDS + w], ub | m);
                    }
                }
                if (colour < 0) colour = c;  // cannot happen: the old colouring is a witness
                tmp_phase[i] = colour;
            }
            gb.sync();
        }
        for (size_t i = tid; i < n; i += stride)
            if (it_type[i] != NB2_ITEM_INVALID) phase[i] = tmp_phase[i];
        gb.sync();
    }
    // ---- Kempe chains: empty the last colour class while it is small (see kempe_plan above)
    // Refinement steps only (an unchanged conflict graph, like the iterated-greedy passes): on the 100k pile a
    // fresh colouring leaves 3718 groups in the last class and the interchanges cost 0.45 ms, one step later the
    // class holds a few hundred.
    for (int sweep = 0; col_edge != nullptr && refine && sweep < 2; ++sweep) {
        __shared__ unsigned int s_kC, s_kN, s_kD;
        int* const max_degree = col_edge + nb * (size_t)NB2_KEMPE_K;  // one spare word behind the table
        for (size_t i = tid; i < NB2_MAX_COLOURS; i += stride) bal[i] = 0u;
        if (tid == 0) *max_degree = 0;
        gb.sync();
        for (size_t i = tid; i < n; i += stride)
            if (it_type[i] != NB2_ITEM_INVALID) atomicAdd(&bal[min((unsigned int)phase[i], (unsigned int)NB2_MAX_COLOURS - 1u)], 1u);
        {  // the largest number of groups on one body (= colours in its mask: the colouring is proper)
            int d = 0;
            for (size_t x = tid; x < nb; x += stride) {
                int dx = 0;
#pragma unroll
                for (int w = 0; w < NB2_MASK_WORDS; ++w) dx += __popcll(__ldcg(&cmask[x * NB2_MASK_WORDS + w]));
                d = max(d, dx);
            }
            d = __reduce_max_sync(0xffffffffu, d);
            if ((threadIdx.x & 31) == 0 && d) atomicMax(max_degree, d);
        }
        gb.sync();
        if (threadIdx.x == 0) {
            unsigned int C = 0;
            for (unsigned int c = 0; c < NB2_MAX_COLOURS; ++c)
                if (__ldcg(&bal[c])) C = c + 1;
            s_kC = C;
            s_kN = C ? __ldcg(&bal[C - 1]) : 0u;
            s_kD = (unsigned int)__ldcg(max_degree);
        }
        __syncthreads();
        gb.sync();  // every block has read the counts before anybody moves on and clears them (a block that saw half-cleared counts would decide differently and the grid barrier would never fill)
        const unsigned int C = s_kC, L = C - 1u;
        // Down to Vizing's bound (largest body degree + 1) and no further: below it the interchanges have to lay
        // alternating colours along the very stacks whose support has to travel through them within a sweep (the
        // 50 x 200 wall went from 5 to 4 colours and its worst penetration at step 120 from 58 to 107 mm).
        if (C < 3u || C > NB2_KEMPE_K || s_kN > NB2_KEMPE_MAX || C <= s_kD + 1u) break;  // uniform over the grid
        // per body and colour: the group that holds the colour there; pending colour changes travel in tmp_phase
        // (a group's colour is only ever read by its home thread)
        for (size_t w = tid; w < nb * (size_t)NB2_KEMPE_K; w += stride) col_edge[w] = -1;
        for (size_t i = tid; i < n; i += stride) tmp_phase[i] = -1;
        gb.sync();
        for (size_t i = tid; i < n; i += stride) {
            if (it_type[i] == NB2_ITEM_INVALID) continue;
            const int c = phase[i], a = it_a[i], b = it_b[i];
            if (a >= 0) __stcg(&col_edge[(size_t)a * NB2_KEMPE_K + c], (int)i);
            if (b >= 0) __stcg(&col_edge[(size_t)b * NB2_KEMPE_K + c], (int)i);
        }
        gb.sync();
        bool emptied = false;
        for (unsigned int r = 0; r < NB2_KEMPE_ROUNDS; ++r) {
            const unsigned int round = 0x400u + (unsigned int)sweep * NB2_KEMPE_ROUNDS + r;  // above the colouring rounds, below the balancing ones
            // bids
            for (size_t i = tid; i < n; i += stride) {
                if (it_type[i] == NB2_ITEM_INVALID || phase[i] != (int)L) continue;
                const unsigned long long prio = ((unsigned long long)round << 52) |
                                                ((unsigned long long)(hash_u32((unsigned int)i + round) & 0xFFFFFu) << 32) |
                                                (unsigned long long)(unsigned int)i;
                kempe_plan((int)i, r, L, it_a, it_b, cmask, col_edge, [&](int x) { atomicMax(&best[x], prio); });
            }
            gb.sync();
            // winners: every body of the plan still carries the bid
            for (size_t i = tid; i < n; i += stride) {
                if (it_type[i] == NB2_ITEM_INVALID || phase[i] != (int)L) continue;
                const unsigned long long prio = ((unsigned long long)round << 52) |
                                                ((unsigned long long)(hash_u32((unsigned int)i + round) & 0xFFFFFu) << 32) |
                                                (unsigned long long)(unsigned int)i;
                bool mine = true;
                const KempePlan P = kempe_plan((int)i, r, L, it_a, it_b, cmask, col_edge,
                                               [&](int x) { mine = mine && __ldcg(&best[x]) == prio; });
                if (P.kind == 0 || !mine) continue;
                const unsigned long long bitL = 1ull << L, bitA = 1ull << P.a;
                if (P.kind == 2) {
                    const unsigned long long bitB = 1ull << P.b;
                    // Along the path every body swaps its a-group and its b-group (the start has only an a-group,
                    // the far end only one of the two, the bodies in between both): swap the two table entries and
                    // the two mask bits of each body right after reading the group that leads on.
                    auto swap_at = [&](int x) {
                        int* ta = &col_edge[(size_t)x * NB2_KEMPE_K + P.a];
                        int* tb = &col_edge[(size_t)x * NB2_KEMPE_K + P.b];
                        const int fa_ = __ldcg(ta), fb_ = __ldcg(tb);
                        __stcg(ta, fb_);
                        __stcg(tb, fa_);
                        const unsigned long long m = __ldcg(&cmask[(size_t)x * NB2_MASK_WORDS]);
                        const unsigned long long sw = (m & ~(bitA | bitB)) | ((m & bitA) ? bitB : 0ull) | ((m & bitB) ? bitA : 0ull);
                        __stcg(&cmask[(size_t)x * NB2_MASK_WORDS], sw);
                    };
                    int x = P.v, want = P.a;
                    for (int k = 0; k < P.n; ++k) {
                        const int f = __ldcg(&col_edge[(size_t)x * NB2_KEMPE_K + want]);
                        const int fa = it_a[f], fb = it_b[f];
                        const int y = fa == x ? fb : fa;
                        __stcg(&tmp_phase[f], want == P.a ? P.b : P.a);  // picked up by the group's home thread
                        swap_at(x);
                        x = y;
                        want = want == P.a ? P.b : P.a;
                    }
                    if (x >= 0) swap_at(x);  // the far end (a static side has no table)
                    const unsigned long long mv = __ldcg(&cmask[(size_t)P.v * NB2_MASK_WORDS]);
                    __stcg(&cmask[(size_t)P.v * NB2_MASK_WORDS], (mv | bitA) & ~bitL);
                    const unsigned long long mu = __ldcg(&cmask[(size_t)P.u * NB2_MASK_WORDS]);
                    __stcg(&cmask[(size_t)P.u * NB2_MASK_WORDS], (mu | bitA) & ~bitL);
                    __stcg(&col_edge[(size_t)P.u * NB2_KEMPE_K + P.a], (int)i);
                    __stcg(&col_edge[(size_t)P.v * NB2_KEMPE_K + P.a], (int)i);
                    __stcg(&col_edge[(size_t)P.u * NB2_KEMPE_K + L], -1);
                    __stcg(&col_edge[(size_t)P.v * NB2_KEMPE_K + L], -1);
                } else {
                    if (P.u >= 0) {
                        const unsigned long long m = __ldcg(&cmask[(size_t)P.u * NB2_MASK_WORDS]);
                        __stcg(&cmask[(size_t)P.u * NB2_MASK_WORDS], (m | bitA) & ~bitL);
                        __stcg(&col_edge[(size_t)P.u * NB2_KEMPE_K + P.a], (int)i);
                        __stcg(&col_edge[(size_t)P.u * NB2_KEMPE_K + L], -1);
                    }
                    if (P.v >= 0) {
                        const unsigned long long m = __ldcg(&cmask[(size_t)P.v * NB2_MASK_WORDS]);
                        __stcg(&cmask[(size_t)P.v * NB2_MASK_WORDS], (m | bitA) & ~bitL);
                        __stcg(&col_edge[(size_t)P.v * NB2_KEMPE_K + P.a], (int)i);
                        __stcg(&col_edge[(size_t)P.v * NB2_KEMPE_K + L], -1);
                    }
                }
                __stcg(&tmp_phase[i], P.a);
            }
            gb.sync();
            // home threads pick their groups' new colours up; is anything left in the class?
            unsigned int left = 0;
            for (size_t i = tid; i < n; i += stride) {
                if (it_type[i] == NB2_ITEM_INVALID) continue;
                const int c = __ldcg(&tmp_phase[i]);
                if (c >= 0) {
                    phase[i] = c;
                    tmp_phase[i] = -1;
                }
                left |= (unsigned int)(phase[i] == (int)L);
            }
            const unsigned int fi = (unsigned int)sweep * NB2_KEMPE_ROUNDS + r;
            if (__syncthreads_or(left) && threadIdx.x == 0) atomicOr(&flags[fi % 3], 1u);
            gb.sync();
            const unsigned int any = *((volatile unsigned int*)&flags[fi % 3]);
            if (tid == 0) flags[(fi + 2) % 3] = 0;
            if (!any) {
                emptied = true;
                break;
            }
        }
        // leave the flag words clean for whoever uses them next
        gb.sync();
        if (tid == 0) flags[0] = flags[1] = flags[2] = 0;
        gb.sync();
        if (!emptied) break;  // uniform: every block read the same flag
    }
    for (size_t i = tid; i < n; i += stride)
        if (it_type[i] != NB2_ITEM_INVALID) raw_phase[i] = phase[i];
    for (size_t i = tid; i < NB2_MAX_COLOURS; i += stride) bal[i] = 0u;
    gb.sync();
    // ---- balancing.  First-fit colouring fills the low colours to the brim (~N_bodies/2 groups) and
    // leaves a tail of nearly empty ones, but every colour costs the solve kernels a grid barrier plus
    // one group's latency chain however few groups it holds, and an over-full colour makes threads run
    // two groups back to back.  Groups of over-full colours therefore migrate to under-full colours
    // that are free at both of their bodies: a deterministic pseudo-random subset (the colour's excess)
    // bids per round, Jones-Plassmann style, so no two movers of a round share a body.
    __shared__ unsigned int s_cnt[NB2_MAX_COLOURS];
    __shared__ unsigned int s_C, s_T, s_over;
    for (size_t i = tid; i < n; i += stride)
        if (it_type[i] != NB2_ITEM_INVALID) atomicAdd(&bal[min((unsigned int)phase[i], (unsigned int)NB2_MAX_COLOURS - 1u)], 1u);
    gb.sync();
    for (unsigned int r = 0; r < NB2_BALANCE_ROUNDS; ++r) {
        const unsigned int round = 0x800u + r;  // above every colouring round: bids of this stage beat the stale ones
        for (unsigned int c = threadIdx.x; c < NB2_MAX_COLOURS; c += blockDim.x) s_cnt[c] = __ldcg(&bal[c]);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int C = 0, total = 0, mx = 0;
            for (unsigned int c = 0; c < NB2_MAX_COLOURS; ++c) {
                if (s_cnt[c]) C = c + 1;
                total += s_cnt[c];
                mx = max(mx, s_cnt[c]);
            }
            s_C = C;
            s_T = C ? (total + C - 1) / C : 0;
            s_over = C > 1 && mx > s_T + s_T / 32 + 8;
        }
        __syncthreads();
        const unsigned int C = s_C, T = s_T;
        if (!s_over) break;  // uniform over the grid: every block read the same counts
        for (int pass = 0; pass < 2; ++pass) {
            for (size_t i = tid; i < n; i += stride) {
                if (it_type[i] == NB2_ITEM_INVALID) continue;
                const unsigned int c = (unsigned int)phase[i];
                if (c >= C || s_cnt[c] <= T) continue;
                if (hash_u32((unsigned int)i * 0x9E3779B9u + round) % s_cnt[c] >= s_cnt[c] - T) continue;
                const int a = it_a[i], b = it_b[i];
                unsigned long long used[NB2_MASK_WORDS];
#pragma unroll
                for (int w = 0; w < NB2_MASK_WORDS; ++w) {
                    used[w] = 0;
                    if (a >= 0) used[w] |= __ldcg(&cmask[(size_t)a * NB2_MASK_WORDS + w]);
                    if (b >= 0) used[w] |= __ldcg(&cmask[(size_t)b * NB2_MASK_WORDS + w]);
                }
                int target = -1;
                unsigned long long best_score = 0;
                for (unsigned int d = 0; d < C; ++d) {
                    if (s_cnt[d] >= T) continue;
                    unsigned long long word = used[0];
#pragma unroll
                    for (int w = 1; w < NB2_MASK_WORDS; ++w)
                        if ((d >> 6) == (unsigned int)w) word = used[w];
                    if ((word >> (d & 63)) & 1ull) continue;
                    const unsigned long long score =
                        (unsigned long long)(T - s_cnt[d]) * (1024u + (hash_u32((unsigned int)i ^ (d * 0x85EBCA6Bu) ^ (round << 20)) & 1023u));
                    if (score > best_score) {
                        best_score = score;
                        target = (int)d;
                    }
                }
                if (target < 0) continue;
                const unsigned long long prio = ((unsigned long long)round << 52) |
                                                ((unsigned long long)(hash_u32((unsigned int)i + round) & 0xFFFFFu) << 32) |
                                                (unsigned long long)(unsigned int)i;
                if (pass == 0) {
                    if (a >= 0) atomicMax(&best[a], prio);
                    if (b >= 0) atomicMax(&best[b], prio);
                } else {
                    if (a >= 0 && __ldcg(&best[a]) != prio) continue;
                    if (b >= 0 && __ldcg(&best[b]) != prio) continue;
                    // sole mover on both bodies this round: their masks are ours to edit
                    const int sides[2] = {a, b};
                    for (int k = 0; k < 2; ++k) {
                        if (sides[k] < 0) continue;
                        unsigned long long* m = &cmask[(size_t)sides[k] * NB2_MASK_WORDS];
                        __stcg(&m[c >> 6], __ldcg(&m[c >> 6]) & ~(1ull << (c & 63)));
                        __stcg(&m[target >> 6], __ldcg(&m[target >> 6]) | (1ull << (target & 63)));
                    }
                    phase[i] = target;
                    atomicAdd(&bal[target], 1u);
                    atomicSub(&bal[c], 1u);
                }
            }
            gb.sync();
        }
    }
}

// ------------------------------------------------------------------ layout
#define NB2_ROW_BUCKETS 4
#define NB2_SMEM_KEYS 1024
__device__ __forceinline__ unsigned int row_bucket(int nrows) {  // 12+ rows -> 0, 9..11 -> 1, 6..8 -> 2, fewer -> 3
    const int b = (12 - min(nrows, 12) + 2) / 3;
    return (unsigned int)min(b, NB2_ROW_BUCKETS - 1);
}
__global__ void k_phase_hist(const unsigned int* __restrict__ changed, size_t n, const int* __restrict__ it_type,
                             const int* __restrict__ it_nrows, const int* __restrict__ phase,
                             unsigned int* ph_bcnt, unsigned int* ph_R, SchedHeader* hdr, unsigned int max_phases) {
    if (*changed == 0u) return;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n && it_type[i] != NB2_ITEM_INVALID;
    const int z = valid ? (it_nrows[i] | (it_type[i] << 8)) : 0;
    const unsigned int rows_generic = valid && !NB2_Z_IS_COMPACT(z) ? (unsigned int)it_nrows[i] : 0u;
    // slots of a phase are handed out by row-count bucket (k_fill_ginfo), longest groups first: the 32 groups a
    // warp of the solve kernels owns then hold the same number of rows (up to one bucket edge per warp).
    // A few dozen (phase, bucket) counters, the per-phase row maxima and the header words would take every
    // item's atomic: they are gathered per block in shared memory first (41 us -> a few us on 296k groups).
    // A schedule in levels has thousands of phases, little contention per counter, and goes direct.
    __shared__ unsigned int s_cnt[NB2_SMEM_KEYS];
    __shared__ unsigned int s_R[NB2_SMEM_KEYS / NB2_ROW_BUCKETS];
    __shared__ unsigned int s_work, s_np;
    const unsigned int nkeys = max_phases * NB2_ROW_BUCKETS;
    const bool blockwise = nkeys <= NB2_SMEM_KEYS;
    if (blockwise) {
        for (unsigned int k = threadIdx.x; k < nkeys; k += blockDim.x) s_cnt[k] = 0u;
        for (unsigned int k = threadIdx.x; k < max_phases; k += blockDim.x) s_R[k] = 0u;
        if (threadIdx.x == 0) s_work = s_np = 0u;
        __syncthreads();
    }
    {  // rows actually scheduled (hdr->work), against the padded slot count
        const unsigned int rows = __reduce_add_sync(0xffffffffu, rows_generic);
        if ((threadIdx.x & 31) == 0 && rows) atomicAdd(blockwise ? &s_work : &hdr->work, rows);
    }
    if (valid) {
        unsigned int p = (unsigned int)phase[i];
        if (p >= max_phases) {
            p = max_phases - 1;
            atomicOr(&hdr->overflow, 2u);
        }
        const unsigned int key = p * NB2_ROW_BUCKETS + row_bucket(it_nrows[i]);
        // generic row slots (ph_R): joints and reference-order groups; compact contact groups reserve none
        if (blockwise) {
            atomicAdd(&s_cnt[key], 1u);
            atomicMax(&s_R[p], rows_generic);
            atomicMax(&s_np, p + 1);
        } else {
            atomicAdd(&ph_bcnt[key], 1u);
            atomicMax(&ph_R[p], rows_generic);
            atomicMax(&hdr->n_phases, p + 1);
        }
    }
    if (blockwise) {
        __syncthreads();
        for (unsigned int k = threadIdx.x; k < nkeys; k += blockDim.x)
            if (s_cnt[k]) atomicAdd(&ph_bcnt[k], s_cnt[k]);
        for (unsigned int k = threadIdx.x; k < max_phases; k += blockDim.x)
            if (s_R[k]) atomicMax(&ph_R[k], s_R[k]);
        if (threadIdx.x == 0) {
            if (s_work) atomicAdd(&hdr->work, s_work);
            if (s_np) atomicMax(&hdr->n_phases, s_np);
        }
    }
}
__global__ void k_phase_scan(const unsigned int* __restrict__ changed, unsigned int* ph_count, unsigned int* ph_bcnt,
                             unsigned int* ph_R, unsigned int* ph_gbase, unsigned int* ph_rbase, SchedHeader* hdr) {
    if (*changed == 0u) return;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned int g = 0, r = 0, mx = 0;
    unsigned int np = hdr->n_phases;
    for (unsigned int p = 0; p < np; ++p) {
        unsigned int base = 0;  // bucket counts -> bucket cursors (k_fill_ginfo hands the slots out)
        for (unsigned int b = 0; b < NB2_ROW_BUCKETS; ++b) {
            const unsigned int c = ph_bcnt[p * NB2_ROW_BUCKETS + b];
            ph_bcnt[p * NB2_ROW_BUCKETS + b] = base;
            base += c;
        }
        ph_count[p] = base;
        ph_gbase[p] = g;
        ph_rbase[p] = r;
        g += ph_count[p];
        r += ph_count[p] * ph_R[p];
        mx = max(mx, ph_count[p]);
    }
    hdr->pad[1] = mx;  // the largest phase: the solve kernels size their blocks to it (one group per thread)
    ph_gbase[np] = g;
    ph_rbase[np] = r;
    hdr->n_groups = g;
    hdr->n_slots = r;
}
__global__ void k_fill_ginfo(const unsigned int* __restrict__ changed, size_t n, const int* __restrict__ it_type,
                             const int* __restrict__ it_a, const int* __restrict__ it_b,
                             const int* __restrict__ it_nrows, const int* __restrict__ phase, int* slot,
                             unsigned int* ph_bcnt, const unsigned int* __restrict__ ph_gbase, int4* g_info,
                             unsigned int max_phases) {
    if (*changed == 0u) return;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n && it_type[i] != NB2_ITEM_INVALID;
    __shared__ unsigned int s_cnt[NB2_SMEM_KEYS];
    const unsigned int nkeys = max_phases * NB2_ROW_BUCKETS;
    const bool blockwise = nkeys <= NB2_SMEM_KEYS;
    unsigned int p = 0, key = 0, sl = 0;
    if (valid) {
        p = min((unsigned int)phase[i], max_phases - 1);
        key = p * NB2_ROW_BUCKETS + row_bucket(it_nrows[i]);
    }
    if (blockwise) {  // rank inside the block, then one reservation per block and counter
        for (unsigned int k = threadIdx.x; k < nkeys; k += blockDim.x) s_cnt[k] = 0u;
        __syncthreads();
        if (valid) sl = atomicAdd(&s_cnt[key], 1u);
        __syncthreads();
        for (unsigned int k = threadIdx.x; k < nkeys; k += blockDim.x)
            if (s_cnt[k]) s_cnt[k] = atomicAdd(&ph_bcnt[k], s_cnt[k]);
        __syncthreads();
        if (valid) sl += s_cnt[key];
    } else if (valid) {
        sl = atomicAdd(&ph_bcnt[key], 1u);
    }
    if (!valid) return;
    slot[i] = (int)sl;
    // z packs the row count (low 8 bits) and the item type (bits 8..) so the solve kernels need
    // no second lookup
    g_info[ph_gbase[p] + sl] = make_int4(it_a[i], it_b[i], it_nrows[i] | (it_type[i] << 8), (int)i);
}
__global__ void k_copy_phase(size_t n, const int* __restrict__ level, int* phase) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) phase[i] = level[i];
}

template <typename K>
static int coop_blocks(Context* ctx, K kernel, int* cache) {
    if (*cache > 0) return NB2_OK;
    int per_sm = 0;
    NB2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPB, 0));
    if (per_sm < 1) return set_error(ctx, NB2_ERR_CUDA, "cooperative kernel does not fit on an SM");
    if (per_sm > 4) per_sm = 4;
    *cache = per_sm * ctx->sm_count;
    return NB2_OK;
}

int launch_schedule(Context* ctx, Sched* s, int mode) {
    const size_t n = s->n_items;
    const unsigned int nb = ctx->n_bodies;
    // Schedule cache (coloured mode): the colouring stays valid while the conflict graph -- the
    // (body pair, row count, type) of every group -- is unchanged.  A device flag is set by comparing
    // this step's groups with the previous step's; every schedule kernel returns at once when it is
    // clear, keeping the previous phases, layout and g_info.  No host read-back is involved.
    NB2_TRY(ctx->flags.reserve(ctx, 4));
    unsigned int* changed = ctx->flags.p + 1;
    unsigned int* n_changed = ctx->flags.p + 3;
    const bool comparable = mode == NB2_MODE_COLOURED && ctx->schedule_cache && s->cache_valid && s->cache_n == n &&
                            n > 0 && s->cache_bodies == nb;
    NB2_CUDA(ctx, cudaMemsetAsync(changed, comparable ? 0 : 0xFF, sizeof(unsigned int), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(n_changed, 0, sizeof(unsigned int), ctx->stream));
    if (mode == NB2_MODE_COLOURED && n > 0) {
        NB2_TRY(s->prev_a.reserve(ctx, n));
        NB2_TRY(s->prev_b.reserve(ctx, n));
        NB2_TRY(s->prev_nt.reserve(ctx, n));
        NB2_TRY(s->prev_b1.reserve(ctx, n));
        NB2_TRY(s->prev_b2.reserve(ctx, n));
        NB2_TRY(ctx->cmask.reserve(ctx, (size_t)nb * NB2_MASK_WORDS + 1));
        if (comparable) {
            k_compare_snapshot<<<nblk(n), TPB, 0, ctx->stream>>>(n, s->it_a.p, s->it_b.p, s->it_nrows.p, s->it_type.p,
                                                                  s->it_b1.p, s->it_b2.p, s->prev_a.p, s->prev_b.p,
                                                                  s->prev_nt.p, s->prev_b1.p, s->prev_b2.p, n_changed);
            ctx->launches++;
        }
        k_refine_decide<<<1, 1, 0, ctx->stream>>>(changed, s->hdr.p, n_changed, (unsigned int)n, comparable ? 1 : 0,
                                                  ctx->incremental_colouring ? 1 : 0);
        k_apply_snapshot<<<nblk(n), TPB, 0, ctx->stream>>>(n, s->it_a.p, s->it_b.p, s->it_nrows.p, s->it_type.p, s->it_b1.p,
                                                            s->it_b2.p, s->prev_a.p, s->prev_b.p, s->prev_nt.p, s->prev_b1.p,
                                                            s->prev_b2.p, changed, s->it_phase.p, ctx->cmask.p, s->ph_count.p,
                                                            s->hdr.p);
        ctx->launches += 2;
        s->cache_valid = true;
        s->cache_n = n;
        s->cache_bodies = nb;
    } else {
        s->cache_valid = false;
    }
    CondRegions cr;
    cr.n = 0;
    cond_add(cr, s->hdr.p, offsetof(SchedHeader, refine_left) / 4, 0);
    cond_add(cr, s->ph_count.p, s->max_phases + 1, 0);
    cond_add(cr, s->ph_R.p, s->max_phases + 1, 0);
    cond_add(cr, s->ph_bcnt.p, 4 * (s->max_phases + 1), 0);
    if (n == 0 || mode == NB2_MODE_REFERENCE_ORDER) cond_launch(ctx, changed, cr);
    if (n == 0) {
        k_phase_scan<<<1, 1, 0, ctx->stream>>>(changed, s->ph_count.p, s->ph_bcnt.p, s->ph_R.p, s->ph_gbase.p, s->ph_rbase.p, s->hdr.p);
        ctx->launches++;
        NB2_CUDA(ctx, cudaGetLastError());
        return NB2_OK;
    }
    NB2_TRY(ctx->barrier.reserve(ctx, NB2_BARRIER_WORDS));
    NB2_CUDA(ctx, cudaMemsetAsync(ctx->barrier.p, 0, 8 * sizeof(unsigned int), ctx->stream));
    unsigned int* flags = ctx->barrier.p + 4;
    if (mode == NB2_MODE_REFERENCE_ORDER) {
        NB2_TRY(ctx->deg.reserve(ctx, (size_t)nb + 1));
        DevBuf<unsigned int>& adj_off = (s == &ctx->ps) ? ctx->adj_off_p : ctx->adj_off;
        DevBuf<int>& adj = (s == &ctx->ps) ? ctx->adj_p : ctx->adj;
        NB2_TRY(adj_off.reserve(ctx, (size_t)nb + 2));
        NB2_TRY(ctx->cursor.reserve(ctx, (size_t)nb + 1));
        NB2_TRY(adj.reserve(ctx, 2 * n + 1));
        NB2_TRY(ctx->pred_a.reserve(ctx, n));
        NB2_TRY(ctx->pred_b.reserve(ctx, n));
        NB2_TRY(ctx->level.reserve(ctx, n));
        NB2_CUDA(ctx, cudaMemsetAsync(ctx->deg.p, 0, ((size_t)nb + 1) * sizeof(unsigned int), ctx->stream));
        NB2_CUDA(ctx, cudaMemsetAsync(ctx->cursor.p, 0, ((size_t)nb + 1) * sizeof(unsigned int), ctx->stream));
        k_degree<<<nblk(n), TPB, 0, ctx->stream>>>(s->it_a.p, s->it_b.p, s->it_type.p, n, ctx->deg.p);
        ctx->launches++;
        NB2_TRY(exclusive_scan_u32(ctx, ctx->deg.p, adj_off.p, nb));
        k_fill_adj<<<nblk(n), TPB, 0, ctx->stream>>>(s->it_a.p, s->it_b.p, s->it_type.p, n, adj_off.p,
                                                     ctx->cursor.p, adj.p);
        k_fill_int<<<nblk(n), TPB, 0, ctx->stream>>>(ctx->pred_a.p, n, -1);
        k_fill_int<<<nblk(n), TPB, 0, ctx->stream>>>(ctx->pred_b.p, n, -1);
        k_fill_int<<<nblk(n), TPB, 0, ctx->stream>>>(ctx->level.p, n, 0);
        k_sort_adj_link<<<nblk(nb), TPB, 0, ctx->stream>>>(nb, adj_off.p, adj.p, s->it_key.p, s->it_a.p,
                                                           ctx->pred_a.p, ctx->pred_b.p);
        ctx->launches += 5;
        NB2_TRY(coop_blocks(ctx, k_levelise, &ctx->coop_blocks_level));
        int blocks = (int)min((size_t)ctx->coop_blocks_level, (n + TPB - 1) / TPB);
        size_t n_ = n;
        const int* ty = s->it_type.p;
        const int* pa = ctx->pred_a.p;
        const int* pb = ctx->pred_b.p;
        int* lv = ctx->level.p;
        unsigned int* bar = ctx->barrier.p;
        void* args[] = {&n_, &ty, &pa, &pb, &lv, &flags, &bar};
        NB2_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_levelise, dim3(blocks), dim3(TPB), args, 0, ctx->stream));
        k_copy_phase<<<nblk(n), TPB, 0, ctx->stream>>>(n, ctx->level.p, s->it_phase.p);
        ctx->launches += 2;
    } else {
        NB2_TRY(ctx->cmask.reserve(ctx, (size_t)nb * NB2_MASK_WORDS + 1));
        NB2_TRY(ctx->best.reserve(ctx, (size_t)nb + 1));
        NB2_TRY(ctx->bal.reserve(ctx, NB2_MAX_COLOURS));
        cond_add(cr, ctx->cmask.p, (size_t)nb * NB2_MASK_WORDS * 2, 1);
        cond_add(cr, ctx->best.p, (size_t)nb * 2, 0);
        cond_add(cr, s->it_phase.p, n, 2);
        cond_add(cr, ctx->bal.p, NB2_MAX_COLOURS, 0);
        cond_launch(ctx, changed, cr);
        NB2_TRY(coop_blocks(ctx, k_colour, &ctx->coop_blocks_colour));
        int blocks = (int)min((size_t)ctx->coop_blocks_colour, (n + TPB - 1) / TPB);
        size_t n_ = n;
        const int* ia = s->it_a.p;
        const int* ib = s->it_b.p;
        const int* ty = s->it_type.p;
        int* ph = s->it_phase.p;
        unsigned long long* cm = ctx->cmask.p;
        unsigned long long* be = ctx->best.p;
        SchedHeader* hd = s->hdr.p;
        unsigned int* bar = ctx->barrier.p;
        const unsigned int* ch = changed;
        unsigned int* bl = ctx->bal.p;
        int* tmp = s->it_slot.p;  // free until k_phase_hist fills it
        size_t nb_ = nb;
        int igp = NB2_IG_PASSES;
        NB2_TRY(s->it_phase_raw.reserve(ctx, n));
        int* raw = s->it_phase_raw.p;
        int* col_edge = nullptr;  // Kempe chains (NB2_KEMPE=0 switches the stage off)
        if (ctx->kempe) {
            NB2_TRY(ctx->col_edge.reserve(ctx, (size_t)nb * NB2_KEMPE_K + 1));
            col_edge = ctx->col_edge.p;
        }
        void* args[] = {&n_, &ia, &ib, &ty, &ph, &cm, &be, &flags, &hd, &bar, &ch, &bl, &tmp, &nb_, &igp, &raw, &col_edge};
        NB2_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_colour, dim3(blocks), dim3(TPB), args, 0, ctx->stream));
        ctx->launches++;
    }
    k_phase_hist<<<nblk(n), TPB, 0, ctx->stream>>>(changed, n, s->it_type.p, s->it_nrows.p, s->it_phase.p,
                                                   s->ph_bcnt.p, s->ph_R.p, s->hdr.p, (unsigned int)s->max_phases);
    k_phase_scan<<<1, 1, 0, ctx->stream>>>(changed, s->ph_count.p, s->ph_bcnt.p, s->ph_R.p, s->ph_gbase.p, s->ph_rbase.p, s->hdr.p);
    k_fill_ginfo<<<nblk(n), TPB, 0, ctx->stream>>>(changed, n, s->it_type.p, s->it_a.p, s->it_b.p, s->it_nrows.p,
                                                   s->it_phase.p, s->it_slot.p, s->ph_bcnt.p, s->ph_gbase.p, s->g_info.p,
                                                   (unsigned int)s->max_phases);
    ctx->launches += 3;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

}  // namespace nb2
