// The two projected Gauss-Seidel loops as persistent cooperative kernels.
//
// k_velocity_solve replaces SORProx::solve (src/solver/sor_prox.rs:48-436): warm start, then
// max_velocity_iterations sweeps.  Each sweep walks the phases of the schedule in order with a
// grid barrier between phases; inside a phase one thread owns one group (rows sharing a body
// pair), keeps the two bodies' mj_lambda in registers and streams its rows from the ELL planes.
// k_position_solve replaces NonlinearSORProx::solve (src/solver/nonlinear_sor_prox.rs:17-310)
// with ncollide's ContactKinematic::contact restated for Plane/Point, Point/Plane, Point/Point.
#include <stdlib.h>

#include "solve_compact.cuh"
#include "solve_position.cuh"

namespace nb2 {

static const int TPB = SOLVE_TPB;

int launch_velocity_solve_coloured(Context* ctx, const SchedDev& sd, const Rows& R, const CompactArrays& CA);
int launch_velocity_solve_staged(Context* ctx, const SchedDev& sd, const Rows& R, int tpb, int depth, int blocks);
int launch_velocity_solve_bulk(Context* ctx, const SchedDev& sd, const Rows& R, int tpb, int depth, int blocks);
int launch_velocity_solve_lockstep(Context* ctx, const SchedDev& sd, const Rows& R, int tpb, int depth, int blocks);
bool staged_geometry(Context* ctx, int* tpb, int* depth, int* blocks);
int launch_position_solve_staged(Context* ctx, const SchedDev& sd, const PosArrays& A, const PosParams& P, int rows_div,
                                 int tpb, int blocks);

// All rows of one group on register-resident mj_lambda (SORProx row updates, sor_prox.rs:181-343;
// warm start :345-435 when `warm`).  Rows are software-pipelined one ahead.
__device__ __forceinline__ void velocity_group(const Rows& R, float4* lam, int4 info, size_t rbase, unsigned int cnt,
                                               size_t g, bool warm) {
    const bool a = info.x >= 0, b = info.y >= 0;
    const int nrows = info.z & 0xFF, type = info.z >> 8;
    RowPkt cur, nxt;
    if (nrows > 0) load_pkt(R, rbase + g, a, b, &cur);
    // Coloured contact groups hold their own normal rows (rows 2*ncc..3*ncc): their impulses are
    // fetched up front so friction rows never chase a dependent load.
    const int ncc = type == NB2_ITEM_CONTACTS ? nrows / 3 : 0;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
    if (ncc > 0) n0 = __ldcg(&R.imp[rbase + (size_t)(2 * ncc + 0) * cnt + g]);
    if (ncc > 1) n1 = __ldcg(&R.imp[rbase + (size_t)(2 * ncc + 1) * cnt + g]);
    if (ncc > 2) n2 = __ldcg(&R.imp[rbase + (size_t)(2 * ncc + 2) * cnt + g]);
    if (ncc > 3) n3 = __ldcg(&R.imp[rbase + (size_t)(2 * ncc + 3) * cnt + g]);
    Lam la, lb;
    la.im = lb.im = 0.f;
    if (a) la = load_lam(lam, info.x);
    if (b) lb = load_lam(lam, info.y);
    for (int r = 0; r < nrows; ++r) {
        const size_t slot = rbase + (size_t)r * cnt + g;
        if (r + 1 < nrows) load_pkt(R, slot + cnt, a, b, &nxt);
        if (cur.kind() != NB2_ROW_NONE) {
            RowJ J;
            unpack_pkt(cur, a, b, la.im, lb.im, &J);
            if (warm) {
                if (cur.imp != 0.f) {
                    if (a) axpy6(cur.imp, J.W1, la.v);
                    if (b) axpy6(cur.imp, J.W2, lb.v);
                }
            } else {
                float dep = 0.f;
                if (cur.kind() == NB2_ROW_DEPENDENT) {
                    if (r < 2 * ncc) {
                        const int k = r >> 1;
                        dep = k == 0 ? n0 : (k == 1 ? n1 : (k == 2 ? n2 : n3));
                    } else {
                        dep = __ldcg(&R.imp[cur.dep()]);  // reference order: another group's row
                    }
                }
                float ni = solve_row(cur.kind(), cur.h, cur.imp, dep, J, a, b, &la, &lb);
                if (ni != cur.imp) __stcg(&R.imp[slot], ni);
            }
        }
        cur = nxt;
    }
    if (a) store_lam(lam, info.x, la);
    if (b) store_lam(lam, info.y, lb);
}

// Phase-by-phase execution with a grid barrier between phases (reference-order levels; also the
// fallback of the coloured mode).  mode_warm: 1 = warm-start pass over the phases first.
__global__ void __launch_bounds__(TPB) k_velocity_solve(SchedDev sd, Rows R, float4* lam, int iters, int mode_warm,
                                                        int symmetric, unsigned int* barrier) {
    GridBarrier gb;
    gb.init(barrier);
    const unsigned int np = sd.hdr->n_phases;
    const size_t tid = interleaved_tid();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = mode_warm ? -1 : 0; it < iters; ++it) {
        for (unsigned int pp = 0; pp < np; ++pp) {
            const unsigned int p = (symmetric && (it & 1)) ? (np - 1 - pp) : pp;
            const unsigned int cnt = sd.ph_count[p];
            const size_t rbase = sd.ph_rbase[p], gbase = sd.ph_gbase[p];
            for (size_t g = tid; g < cnt; g += stride)
                velocity_group(R, lam, sd.g_info[gbase + g], rbase, cnt, g, it < 0);
            gb.sync();
        }
    }
}

// Reference-order warm start: one thread per body accumulates the impulses of its incident rows
// in the reference's warm-start order (contact unilateral, unilateral_ground, bilateral,
// bilateral_ground, then the joint buckets; sor_prox.rs:19-45,57-58).
__global__ void __launch_bounds__(TPB) k_warmstart_ref(unsigned int nb, const unsigned int* __restrict__ adj_off,
                                                       const int* __restrict__ adj, SchedDev sd, Rows R,
                                                       const int* __restrict__ it_nrows, float4* lam) {
    unsigned int body = blockIdx.x * blockDim.x + threadIdx.x;
    if (body >= nb) return;
    const unsigned int s = adj_off[body], e = adj_off[body + 1];
    if (s == e) return;
    Lam l;
#pragma unroll
    for (int k = 0; k < 6; ++k) l.v[k] = 0.f;
    l.im = lam[2 * body].w;
    const int order[6] = {4, 5, 2, 3, 0, 1};
    for (int ob = 0; ob < 6; ++ob) {
        const unsigned long long bucket = (unsigned long long)order[ob];
        for (unsigned int k = s; k < e; ++k) {
            const int item = adj[k];
            if ((sd.it_key[item] >> 40) != bucket) continue;
            const unsigned int p = min((unsigned int)sd.it_phase[item], sd.max_phases - 1);
            const unsigned int cnt = sd.ph_count[p];
            const size_t rbase = sd.ph_rbase[p];
            const bool side_a = sd.it_a[item] == (int)body;
            const int nrows = it_nrows[item];
            for (int r = 0; r < nrows; ++r) {
                const size_t slot = rbase + (size_t)r * cnt + (size_t)sd.it_slot[item];
                if (row_kind(R, slot) == NB2_ROW_NONE) continue;
                const float impulse = R.imp[slot];
                if (impulse == 0.f) continue;
                RowJ J;
                load_row_j(R, slot, side_a, !side_a, l.im, l.im, &J);
                axpy6(impulse, side_a ? J.W1 : J.W2, l.v);
            }
        }
    }
    store_lam(lam, (int)body, l);
}

// ------------------------------------------------------------------------------------------
// position solve
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB) k_position_solve(SchedDev sd, PosArrays A, const nb2_joint* __restrict__ joints,
                                                        const nb2_manifold* __restrict__ manifolds,
                                                        const unsigned int* __restrict__ chunk_manifold,
                                                        const float4* __restrict__ p_row, size_t P_stride, PosParams P,
                                                        int iters, int rows_div, unsigned int* barrier) {
    GridBarrier gb;
    gb.init(barrier);
    const unsigned int np = sd.hdr->n_phases;
    const size_t tid = interleaved_tid();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        for (unsigned int p = 0; p < np; ++p) {
            const unsigned int cnt = sd.ph_count[p];
            const size_t gbase = sd.ph_gbase[p];
            for (size_t g = tid; g < cnt; g += stride) {
                const int4 info = sd.g_info[gbase + g];
                const int item = info.w;
                const int type = sd.it_type[item];
                const int src = sd.it_src[item];
                PosBody b1, b2;
                if (type == NB2_ITEM_JOINT) {
                    const nb2_joint& j = joints[src];
                    load_pos_body(A, j.body1, &b1);
                    load_pos_body(A, j.body2, &b2);
                    joint_position(j, &b1, &b2, P);
                    if (b1.dynamic) store_pos_body(A, j.body1, b1);
                    if (b2.dynamic) store_pos_body(A, j.body2, b2);
                } else {
                    const unsigned int m = chunk_manifold[src];
                    const nb2_manifold& mf = manifolds[m];
                    load_pos_body(A, mf.body1, &b1);
                    load_pos_body(A, mf.body2, &b2);
                    const Pose c1 = load_coll(mf.coll1_wrt_body), c2 = load_coll(mf.coll2_wrt_body);
                    const int nrows = (info.z & 0xFF) / rows_div;
                    bool moved1 = false, moved2 = false;
                    for (int lcc = 0; lcc < nrows; ++lcc) {
                        const size_t ps = (size_t)NB2_CHUNK * gbase + (size_t)lcc * cnt + g;
                        const float4 l1 = __ldg(&p_row[0 * P_stride + ps]);
                        const float4 l2 = __ldg(&p_row[1 * P_stride + ps]);
                        const float4 d1 = __ldg(&p_row[2 * P_stride + ps]);
                        const float4 d2 = __ldg(&p_row[3 * P_stride + ps]);
                        const float4 n1 = __ldg(&p_row[4 * P_stride + ps]);
                        // update_contact_constraint (nonlinear_sor_prox.rs:156-294)
                        const Pose m1 = pose_mul(b1.bp.pose, c1), m2 = pose_mul(b2.bp.pose, c2);
                        ContactEval ce;
                        if (!kinematic_contact(l1, l2, d1, d2, n1, m1, m2, &ce)) continue;
                        const float rhs = clamp_rhs(-ce.depth, false, P);
                        if (rhs >= 0.f) continue;
                        Vec3 w1l = mk3(0.f, 0.f, 0.f), w1a = w1l, w2l = w1l, w2a = w1l;
                        float inv_r = 0.f;
                        pos_fill(b1, ce.world1, false, -ce.normal, &w1l, &w1a, &inv_r);
                        pos_fill(b2, ce.world2, false, ce.normal, &w2l, &w2a, &inv_r);
                        if (inv_r == 0.f) continue;
                        const float r = 1.f / inv_r;
                        const float impulse = -rhs * r;  // solve_unilateral, :137-152
                        if (b1.dynamic) {
                            apply_displacement(&b1.bp, b1.local_com, w1l * impulse, w1a * impulse);
                            moved1 = true;
                        }
                        if (b2.dynamic) {
                            apply_displacement(&b2.bp, b2.local_com, w2l * impulse, w2a * impulse);
                            moved2 = true;
                        }
                    }
                    if (moved1) store_pos_body(A, mf.body1, b1);
                    if (moved2) store_pos_body(A, mf.body2, b2);
                }
            }
            gb.sync();
        }
    }
}

// ------------------------------------------------------------------------------------------
// diagnostics: natural-map residual of the velocity rows, max penetration
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    atomicMax((int*)addr, __float_as_int(v));  // valid for v >= 0
}
// order-preserving float -> int key (any sign); key 0x80000000 = "no value"
__device__ __forceinline__ int float_key(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : (i ^ 0x7FFFFFFF);
}
__global__ void __launch_bounds__(TPB) k_residual(SchedDev sd, Rows R, CompactArrays CA,
                                                  const float4* __restrict__ lam, float* res_max, double* res_sq,
                                                  unsigned int* res_n, unsigned int* rows_two,
                                                  unsigned int* rows_ground) {
    const unsigned int np = sd.hdr->n_phases;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float mx = 0.f;
    double sq = 0.0;
    unsigned int cntr = 0, two = 0, ground = 0;
    for (unsigned int p = 0; p < np; ++p) {
        const unsigned int cnt = sd.ph_count[p];
        const size_t rbase = sd.ph_rbase[p], gbase = sd.ph_gbase[p];
        for (size_t g = tid; g < cnt; g += stride) {
            const int4 info = sd.g_info[gbase + g];
            const bool a = info.x >= 0, b = info.y >= 0;
            if (NB2_Z_IS_COMPACT(info.z) && CA.c_geo != nullptr) {
                unsigned int nr = 0;
                compact_group(CA, 2, info.x, info.y, (size_t)NB2_CHUNK * gbase, cnt, g, (info.z & 0xFF) >> 4, &mx, &sq, &nr);
                cntr += nr;
                if (a && b) two += nr; else ground += nr;
                continue;
            }
            Lam la, lb;
            la.im = lb.im = 0.f;
            if (a) la = load_lam(lam, info.x);
            if (b) lb = load_lam(lam, info.y);
            for (int r = 0; r < (info.z & 0xFF); ++r) {
                const size_t slot = rbase + (size_t)r * cnt + g;
                const float4 q4 = R.jac[4 * R.S + slot];
                const int2 meta = make_int2(__float_as_int(q4.z), __float_as_int(q4.w));
                if (meta.x == NB2_ROW_NONE) continue;
                const float impulse = R.imp[slot];
                RowJ J;
                load_row_j(R, slot, a, b, la.im, lb.im, &J);
                const float4 h = R.hdr[slot];
                float lo = 0.f, hi = NB2_F32_MAX;
                if (meta.x == NB2_ROW_BILATERAL) {
                    lo = h.z;
                    hi = h.w;
                } else if (meta.x == NB2_ROW_DEPENDENT) {
                    hi = h.z * R.imp[meta.y];
                    lo = -hi;
                }
                float w = (a ? dot6(J.J1, la.v) : 0.f) + (b ? dot6(J.J2, lb.v) : 0.f) + h.x;
                float ni = meta.x == NB2_ROW_UNILATERAL ? fmaxf(impulse - h.y * w, 0.f)
                                                        : clampf(impulse - h.y * w, lo, hi);
                float d = fabsf(ni - impulse);
                mx = fmaxf(mx, d);
                sq += (double)d * d;
                ++cntr;
                if (a && b) ++two; else ++ground;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
        sq += __shfl_down_sync(0xffffffffu, sq, o);
        cntr += __shfl_down_sync(0xffffffffu, cntr, o);
        two += __shfl_down_sync(0xffffffffu, two, o);
        ground += __shfl_down_sync(0xffffffffu, ground, o);
    }
    if ((threadIdx.x & 31) == 0 && cntr) {
        atomic_max_nonneg(res_max, mx);
        atomicAdd(res_sq, sq);
        atomicAdd(res_n, cntr);
        atomicAdd(rows_two, two);
        atomicAdd(rows_ground, ground);
    }
}
// max penetration at the current poses (order-preserving int key, see float_key)
__global__ void __launch_bounds__(TPB) k_penetration(SchedDev sd, PosArrays A,
                                                     const nb2_manifold* __restrict__ manifolds,
                                                     const unsigned int* __restrict__ chunk_manifold,
                                                     const float4* __restrict__ p_row, size_t P_stride,
                                                     int rows_div, int* pen_max) {
    const unsigned int np = sd.hdr->n_phases;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int mx = (int)0x80000000;
    for (unsigned int p = 0; p < np; ++p) {
        const unsigned int cnt = sd.ph_count[p];
        const size_t gbase = sd.ph_gbase[p];
        for (size_t g = tid; g < cnt; g += stride) {
            const int4 info = sd.g_info[gbase + g];
            const int item = info.w;
            if (sd.it_type[item] == NB2_ITEM_JOINT) continue;
            const unsigned int m = chunk_manifold[sd.it_src[item]];
            const nb2_manifold& mf = manifolds[m];
            PosBody b1, b2;
            load_pos_body(A, mf.body1, &b1);
            load_pos_body(A, mf.body2, &b2);
            const Pose m1 = pose_mul(b1.bp.pose, load_coll(mf.coll1_wrt_body));
            const Pose m2 = pose_mul(b2.bp.pose, load_coll(mf.coll2_wrt_body));
            const int nrows = (info.z & 0xFF) / rows_div;
            for (int lcc = 0; lcc < nrows; ++lcc) {
                const size_t ps = (size_t)NB2_CHUNK * gbase + (size_t)lcc * cnt + g;
                ContactEval ce;
                if (kinematic_contact(p_row[0 * P_stride + ps], p_row[1 * P_stride + ps], p_row[2 * P_stride + ps],
                                      p_row[3 * P_stride + ps], p_row[4 * P_stride + ps], m1, m2, &ce))
                    mx = max(mx, float_key(ce.depth));
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx != (int)0x80000000) atomicMax(pen_max, mx);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static Rows rows_of(Context* ctx) {
    Rows R;
    R.jac = ctx->r_jac.p;
    R.hdr = ctx->r_hdr.p;
    R.imp = ctx->r_imp.p;
    R.S = ctx->n_slots_max;
    return R;
}
static CompactArrays compact_arrays(Context* ctx, bool ref) {
    CompactArrays A;
    A.raw = ctx->raw.p;
    A.com_im = ctx->com_im.p;
    A.inv_i = ctx->inv_i.p;
    A.lam = ctx->lam.p;
    A.c_geo = ctx->step_layout == 1 ? ctx->c_geo.p : nullptr;
    A.P = ctx->n_pslots_max;
    A.any_mask = ctx->any_mask ? 1 : 0;
    return A;
}
static PosArrays pos_arrays(Context* ctx) {
    PosArrays A;
    A.raw = ctx->raw.p;
    A.pos_t = ctx->pos_t.p;
    A.pos_q = ctx->pos_q.p;
    A.com_im = ctx->com_im.p;
    A.inv_i = ctx->inv_i.p;
    return A;
}
template <typename K>
static int coop_limit(Context* ctx, K kernel, int* cache) {
    if (*cache > 0) return NB2_OK;
    int per_sm = 0;
    NB2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPB, 0));
    if (per_sm < 1) return set_error(ctx, NB2_ERR_CUDA, "cooperative kernel does not fit on an SM");
    if (per_sm > 8) per_sm = 8;
    *cache = per_sm * ctx->sm_count;
    return NB2_OK;
}

int launch_velocity_solve(Context* ctx, int mode) {
    const bool ref = mode == NB2_MODE_REFERENCE_ORDER;
    if (ctx->vs.n_items == 0) return NB2_OK;
    SchedDev sd = sched_dev(ctx->vs);
    Rows R = rows_of(ctx);
    if (ref) {
        unsigned int nb = ctx->n_bodies;
        k_warmstart_ref<<<(nb + TPB - 1) / TPB, TPB, 0, ctx->stream>>>(nb, ctx->adj_off.p, ctx->adj.p, sd, R,
                                                                       ctx->vs.it_nrows.p, ctx->lam.p);
        ctx->launches++;
    }
    NB2_TRY(coop_limit(ctx, k_velocity_solve, &ctx->coop_blocks_vel));
    float4* lam = ctx->lam.p;
    int iters = (int)ctx->params.max_velocity_iterations;
    int warm = ref ? 0 : 1;
    unsigned int* bar = ctx->barrier.p + NB2_BARRIER_VELOCITY;  // zeroed once per step (api.cu)
    size_t want = (ctx->vs.n_items + TPB - 1) / TPB;
    int blocks = (int)(want < (size_t)ctx->coop_blocks_vel ? want : (size_t)ctx->coop_blocks_vel);
    if (blocks < 1) blocks = 1;
    if (ref && ctx->ref_blocks > 0 && ctx->ref_blocks < blocks) blocks = ctx->ref_blocks;
    // alternating the colour order per sweep (symmetric Gauss-Seidel) was measured: it helps flat
    // piles and hurts tall ones (profiles/r01_notes.md), so the plain order stays the default
    int symmetric = 0;
    if (ctx->step_layout == 1) return launch_velocity_solve_coloured(ctx, sd, R, compact_arrays(ctx, ref));
    if (!ref && ctx->velocity_kernel >= 2) {
        int tpb_s, depth_s, blocks_s;
        if (staged_geometry(ctx, &tpb_s, &depth_s, &blocks_s)) {
            if (ctx->velocity_kernel == 3) return launch_velocity_solve_bulk(ctx, sd, R, tpb_s, depth_s, blocks_s);
            // 2 = the free-running per-thread rings.  The layout hands the slots of a phase out by row count
            // (schedule.cu, k_fill_ginfo), so the 32 groups of a warp hold equal row counts and the warp's copies
            // stay coalesced without walking in lockstep: live 100k pile 0.79 ms (lockstep 0.89 ms, and 1.01 ms
            // before the slots were sorted), 4096 x pyramid3 21.4 ms (24.0 / 25.6).  4 forces the lockstep stream.
            const bool lockstep = ctx->velocity_kernel == 4;
            if (lockstep) return launch_velocity_solve_lockstep(ctx, sd, R, tpb_s, depth_s, blocks_s);
            return launch_velocity_solve_staged(ctx, sd, R, tpb_s, depth_s, blocks_s);
        }
    }
    void* args[] = {&sd, &R, &lam, &iters, &warm, &symmetric, &bar};
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[6], ctx->stream));
    NB2_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_velocity_solve, dim3(blocks), dim3(TPB), args, 0, ctx->stream));
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[7], ctx->stream));
    ctx->launches++;
    return NB2_OK;
}

int launch_position_solve(Context* ctx, int mode) {
    const bool ref = mode == NB2_MODE_REFERENCE_ORDER;
    Sched& s = ref ? ctx->ps : ctx->vs;
    if (s.n_items == 0 || ctx->params.max_position_iterations == 0) return NB2_OK;
    SchedDev sd = sched_dev(s);
    PosArrays A = pos_arrays(ctx);
    PosParams P;
    P.erp = ctx->params.erp;
    P.allowed_lin = ctx->params.allowed_linear_error;
    P.allowed_ang = ctx->params.allowed_angular_error;
    P.max_lin = ctx->params.max_linear_correction;
    P.max_ang = ctx->params.max_angular_correction;
    NB2_TRY(coop_limit(ctx, k_position_solve, &ctx->coop_blocks_pos));
    const nb2_joint* joints = ctx->joints.p;
    const nb2_manifold* manifolds = ctx->manifolds.p;
    const unsigned int* cm = ctx->chunk_manifold.p;
    const float4* prow = ctx->p_row.p;
    size_t pstride = ctx->n_pslots_max;
    int iters = (int)ctx->params.max_position_iterations;
    unsigned int* bar = ctx->barrier.p + NB2_BARRIER_POSITION;
    size_t want = (s.n_items + TPB - 1) / TPB;
    int blocks = (int)(want < (size_t)ctx->coop_blocks_pos ? want : (size_t)ctx->coop_blocks_pos);
    if (blocks < 1) blocks = 1;
    if (ref && ctx->ref_blocks > 0 && ctx->ref_blocks < blocks) blocks = ctx->ref_blocks;
    // contacts per group: reference order -> the row count itself; coloured rows -> 3 rows per contact;
    // compact -> the count sits in bits 4..7
    int rows_div = ref ? 1 : (ctx->step_layout == 1 ? 16 : (ctx->contact_model == 1 ? 1 : 3));
    if (!ref && ctx->velocity_kernel >= 2) {
        int tpb_s, depth_s, blocks_s;
        if (staged_geometry(ctx, &tpb_s, &depth_s, &blocks_s)) {
            int rc = launch_position_solve_staged(ctx, sd, A, P, rows_div, tpb_s, blocks_s);
            if (rc != NB2_STAGED_NOT_APPLICABLE) return rc;  // NB2_OK or a real error
        }
    }
    void* args[] = {&sd, &A, &joints, &manifolds, &cm, &prow, &pstride, &P, &iters, &rows_div, &bar};
    NB2_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_position_solve, dim3(blocks), dim3(TPB), args, 0, ctx->stream));
    ctx->launches++;
    return NB2_OK;
}

__global__ void k_count_broken(const nb2_joint* __restrict__ joints, unsigned int n, unsigned int* out) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int v = __ballot_sync(0xffffffffu, i < n && joints[i].broken != 0u);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, (unsigned int)__popc(v));
}

// stat_f layout: [0] res_max ; stat_u[4] pen_max key ; stat_d: [0] res_sq [1] energy ; stat_u: [0] res_n [1] rows_two
// [2] rows_ground [3] non_finite [5] broken joints
int launch_stats(Context* ctx, int mode) {
    const bool ref = mode == NB2_MODE_REFERENCE_ORDER;
    NB2_TRY(ctx->stat_f.reserve(ctx, 16));
    NB2_TRY(ctx->stat_u.reserve(ctx, 16));
    NB2_CUDA(ctx, cudaMemsetAsync(ctx->stat_f.p, 0, 16 * sizeof(float), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(ctx->stat_u.p, 0, 16 * sizeof(unsigned int), ctx->stream));
    {
        const unsigned int none = 0x80000000u;
        NB2_CUDA(ctx, cudaMemcpyAsync(ctx->stat_u.p + 4, &none, sizeof(none), cudaMemcpyHostToDevice, ctx->stream));
    }
    float* f = ctx->stat_f.p;
    double* d = (double*)(ctx->stat_f.p + 4);  // 8-byte aligned inside the float scratch
    unsigned int* u = ctx->stat_u.p;
    const int blocks = ctx->sm_count * 4;
    if (ctx->vs.n_items) {
        k_residual<<<blocks, TPB, 0, ctx->stream>>>(sched_dev(ctx->vs), rows_of(ctx), compact_arrays(ctx, ref), ctx->lam.p,
                                                    f + 0, d + 0, u + 0, u + 1, u + 2);
        ctx->launches++;
        Sched& s = ref ? ctx->ps : ctx->vs;
        if (ctx->n_contacts) {
            k_penetration<<<blocks, TPB, 0, ctx->stream>>>(sched_dev(s), pos_arrays(ctx), ctx->manifolds.p,
                                                           ctx->chunk_manifold.p, ctx->p_row.p, ctx->n_pslots_max,
                                                           ref ? 1 : (ctx->step_layout == 1 ? 16 : (ctx->contact_model == 1 ? 1 : 3)),
                                                           (int*)(u + 4));
            ctx->launches++;
        }
    }
    NB2_TRY(launch_body_stats(ctx, d + 1, u + 3));
    if (ctx->n_joints) {
        k_count_broken<<<(ctx->n_joints + TPB - 1) / TPB, TPB, 0, ctx->stream>>>(ctx->joints.p, ctx->n_joints, u + 5);
        ctx->launches++;
    }
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

}  // namespace nb2
