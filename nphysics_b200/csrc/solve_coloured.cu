// The coloured production velocity kernel.  This translation unit is compiled WITH FMA contraction
// (no -fmad=false): the coloured mode is judged on residual / penetration / energy, not on bit
// equality with the sequential reference order, and half the floating-point instructions matter
// here -- a phase of the 100k-box pile has only ~6 resident warps per SM, so the kernel is bound by
// instruction issue latency, not by memory (profiles/r01_notes.md).
#define NB2_COLOURED_TU 1
#include <stdio.h>
#include <stdlib.h>

#include "solve_compact.cuh"
#include "solve_position.cuh"

namespace nb2 {

// Generic rows of one group (joints in coloured mode): plain sequential loads, rare path.
__device__ __noinline__ void generic_group(const Rows& R, float4* lam, int4 info, size_t rbase, unsigned int cnt, size_t g,
                                           bool warm) {
    const bool a = info.x >= 0, b = info.y >= 0;
    const int nrows = info.z & 0xFF;
    Lam la, lb;
    la.im = lb.im = 0.f;
    if (a) la = load_lam(lam, info.x);
    if (b) lb = load_lam(lam, info.y);
    for (int r = 0; r < nrows; ++r) {
        const size_t slot = rbase + (size_t)r * cnt + g;
        RowPkt k;
        load_pkt(R, slot, a, b, &k);
        if (k.kind() == NB2_ROW_NONE) continue;
        RowJ J;
        unpack_pkt(k, a, b, la.im, lb.im, &J);
        if (warm) {
            if (k.imp != 0.f) {
                if (a) axpy6(k.imp, J.W1, la.v);
                if (b) axpy6(k.imp, J.W2, lb.v);
            }
            continue;
        }
        const float dep = k.kind() == NB2_ROW_DEPENDENT ? __ldcg(&R.imp[k.dep()]) : 0.f;
        const float ni = solve_row(k.kind(), k.h, k.imp, dep, J, a, b, &la, &lb);
        if (ni != k.imp) __stcg(&R.imp[slot], ni);
    }
    if (a) store_lam(lam, info.x, la);
    if (b) store_lam(lam, info.y, lb);
}

// The coloured production kernel: compact contact groups inline, joint groups through generic_group.
__global__ void __launch_bounds__(SOLVE_TPB) k_velocity_solve_coloured(SchedDev sd, Rows R, CompactArrays CA, int iters,
                                                                 unsigned int* barrier) {
    GridBarrier gb;
    gb.init(barrier);
    const unsigned int np = sd.hdr->n_phases;
    const size_t tid = interleaved_tid();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = -1; it < iters; ++it) {
        const int what = it < 0 ? 1 : 0;
        for (unsigned int p = 0; p < np; ++p) {
            const unsigned int cnt = sd.ph_count[p];
            const size_t rbase = sd.ph_rbase[p], gbase = sd.ph_gbase[p];
            for (size_t g = tid; g < cnt; g += stride) {
                const int4 info = __ldg(&sd.g_info[gbase + g]);
                if (NB2_Z_IS_COMPACT(info.z))
                    compact_group(CA, what, info.x, info.y, (size_t)NB2_CHUNK * gbase, cnt, g, (info.z & 0xFF) >> 4, nullptr,
                                  nullptr, nullptr);
                else
                    generic_group(R, CA.lam, info, rbase, cnt, g, it < 0);
            }
            gb.sync();
        }
    }
}


// ------------------------------------------------------------------------------------------
// Staged row stream (the default coloured velocity kernel).
//
// With a barrier per colour the barrier-to-barrier critical path of the register-pipelined kernel is
// a chain of dependent DRAM round trips: g_info -> first row -> ... -> twelfth row.  None of those
// loads depends on what other groups computed: the thread -> group mapping is static, and the row
// planes (J, M^-1 J, rhs/r/limits, kind) are constant during the solve.  So every thread runs a
// PRODUCER D entries ahead of its own consumption, across phase and sweep boundaries: it walks the
// same (sweep, phase, group, row) sequence and copies each row's 6 quads into its private
// slice of a shared-memory ring with cp.async (16 bytes per lane, coalesced over the group index).
// While the block waits at a grid barrier the rows of its next groups are already landing.  After
// the barrier only the mutable data is fetched (the two bodies' mj_lambda and the group's impulses,
// L2 hits, issued together), then the rows are consumed from shared memory.
//
// Ring entry = 7 quads x TPB lanes: planes 0..4 jacobian, 5 header, 6 control.  A group is announced
// by an entry whose control quad holds its g_info (written with st.shared), followed by its rows.  Entries are private to their thread, so cp.async.wait_group is the only
// synchronisation.  A slot is refilled one consume step after it was read, when the values read
// from it have already been used.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(unsigned int dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// row planes are read once per sweep and are far larger than L2: evict-first keeps the small mutable
// state (mj_lambda, impulses, g_info) L2-resident between phases
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_evict_normal_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16_ef(unsigned int dst, const void* src, unsigned long long pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned int dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#ifdef NB2_TRACE  // developer build (make EXTRA=-DNB2_TRACE): per-phase timestamps, see DESIGN.md "Developer knobs"
__device__ unsigned long long g_trace[4096];
__device__ unsigned long long g_ptrace[512 * 8];  // position kernel: per phase visit, see the printf at its end
#endif
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {  // pending in 0..2
    if (pending <= 0) cp_async_wait<0>();
    else if (pending == 1) cp_async_wait<1>();
    else cp_async_wait<2>();
}

struct StreamPos {
    int s;           // sweep, 0 = warm start
    unsigned int p;  // phase
    unsigned int g;  // group within the phase
};
// moves q forward to the first existing group at or after q (thread `tid` owns groups tid, tid+stride, ...)
__device__ __forceinline__ bool seek_group(StreamPos& q, unsigned int tid, unsigned int np, int iters,
                                           const unsigned int* s_cnt) {
    for (;;) {
        if (q.s > iters) return false;
        if (q.g < s_cnt[q.p]) return true;
        q.g = tid;
        if (++q.p == np) {
            q.p = 0;
            ++q.s;
        }
    }
}

// The block size is a launch-time value: the launcher sizes the block to the groups an SM owns per
// phase (one group per thread).  D = ring entries per thread.
// plane base pointers as kernel parameters (constant bank): one IMAD.WIDE per address
struct StagedRows {
    const float4* q[6];  // 5 jacobian planes + header
    float* imp;
};
#define NB2_VENTRY 7
#define NB2_STAGED_MAX_DEPTH 5
#define NB2_STAGED_MIN_DEPTH 3
#define NB2_STAGED_DEPTH 4

template <int D>
__global__ void __launch_bounds__(384, 1) k_velocity_solve_staged(SchedDev sd, StagedRows R, float4* lam, int iters,
                                                                  unsigned int* barrier, int trace) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned int TPBK = blockDim.x;
    float4* ring = reinterpret_cast<float4*>(smem_raw);                    // [D][NB2_VENTRY][TPBK]
    float* simp = reinterpret_cast<float*>(ring + D * NB2_VENTRY * TPBK);  // [12][TPBK] impulses of the running group
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(simp + 12 * TPBK);
    unsigned int* s_rbase = s_cnt + NB2_MAX_COLOURS;
    unsigned int* s_gbase = s_rbase + NB2_MAX_COLOURS;
    const unsigned int np = min(sd.hdr->n_phases, (unsigned int)NB2_MAX_COLOURS);
    if (np == 0) return;
    for (unsigned int i = threadIdx.x; i < np; i += TPBK) {
        s_cnt[i] = sd.ph_count[i];
        s_rbase[i] = sd.ph_rbase[i];
        s_gbase[i] = sd.ph_gbase[i];
    }
    __syncthreads();
    GridBarrier gb;
    gb.init(barrier);
    const unsigned int t = threadIdx.x;
    const unsigned int tid = (unsigned int)interleaved_tid();
    const unsigned int stride = gridDim.x * TPBK;
    const unsigned int ring_u32 = (unsigned int)__cvta_generic_to_shared(ring) + t * 16u;
    const unsigned int plane_b = TPBK * 16u, entry_b = NB2_VENTRY * plane_b;

    // ---- producer state: the group being copied, and the one after it (its g_info already in flight)
    StreamPos pc = {0, 0u, tid}, pn;
    bool vc = seek_group(pc, tid, np, iters, s_cnt), vn = false;
    int4 ic = make_int4(-1, -1, 0, 0), in_ = ic;
    if (vc) {
        ic = __ldg(&sd.g_info[s_gbase[pc.p] + pc.g]);
        pn = pc;
        pn.g += stride;
        vn = seek_group(pn, tid, np, iters, s_cnt);
        if (vn) in_ = __ldg(&sd.g_info[s_gbase[pn.p] + pn.g]);
    }
    int pr = -1;  // -1: the header entry of `pc` comes next
    const unsigned long long pol = (trace & 0x40000000) ? l2_evict_normal_policy() : l2_evict_first_policy();  // bit 30: A/B of the hint
    unsigned int pslot = 0, pcnt = 0;  // row slot being copied and its stride (groups of the phase)
    auto produce = [&](int e) {
        if (vc) {
            const unsigned int dst = ring_u32 + (unsigned int)e * entry_b;
            if (pr < 0) {
                reinterpret_cast<int4*>(ring)[(e * NB2_VENTRY + 6) * TPBK + t] = ic;
                pcnt = s_cnt[pc.p];
                pslot = s_rbase[pc.p] + pc.g;
            } else {
                cp_async16_ef(dst + 1u * plane_b, R.q[1] + pslot, pol);
                cp_async16_ef(dst + 3u * plane_b, R.q[3] + pslot, pol);
                cp_async16_ef(dst + 4u * plane_b, R.q[4] + pslot, pol);
                cp_async16_ef(dst + 5u * plane_b, R.q[5] + pslot, pol);
                cp_async16_ef(dst, R.q[0] + pslot, pol);  // both sides unconditionally: one branch less per row, and the
                cp_async16_ef(dst + 2u * plane_b, R.q[2] + pslot, pol);  // absent side of a ground row is only 32 bytes
                pslot += pcnt;
            }
            if (++pr >= (ic.z & 0xFF)) {
                pc = pn;
                ic = in_;
                vc = vn;
                pr = -1;
                if (vn) {
                    pn.g += stride;
                    vn = seek_group(pn, tid, np, iters, s_cnt);
                    if (vn) in_ = __ldg(&sd.g_info[s_gbase[pn.p] + pn.g]);
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (int e = 0; e < D; ++e) produce(e);

    int ce = 0;        // ring entry consumed next
    bool any = false;  // false until the first entry was consumed (nothing to refill yet)
    auto take = [&]() -> int {  // makes entry `ce` readable and returns it
        cp_async_wait<D - 2>();
        const int e = ce;
        ce = ce + 1 == D ? 0 : ce + 1;
        return e;
    };
    auto refill_prev = [&](int e) {  // refills the entry consumed one step ago (its values are used up by now)
        if (any) produce(e == 0 ? D - 1 : e - 1);
        any = true;
    };
    // next row of the running group: ring -> registers, then the refill of the previous entry
    auto next_row = [&](int r, RowPkt* k) {
        const int e = take();
        const float4* q = ring + (e * NB2_VENTRY) * TPBK + t;
        k->q0 = q[0 * TPBK];
        k->q1 = q[1 * TPBK];
        k->q2 = q[2 * TPBK];
        k->q3 = q[3 * TPBK];
        k->q4 = q[4 * TPBK];
        k->h = q[5 * TPBK];
        k->imp = simp[r * TPBK + t];
        refill_prev(e);
    };

    for (int s = 0; s <= iters; ++s) {  // s == 0: warm start (sor_prox.rs:57-58, 345-435)
        for (unsigned int p = 0; p < np; ++p) {
            const unsigned int cnt = s_cnt[p];
            const unsigned int rbase = s_rbase[p];
            for (unsigned int g = tid; g < cnt; g += stride) {
                const int e0 = take();
                const int4 info = reinterpret_cast<const int4*>(ring)[(e0 * NB2_VENTRY + 6) * TPBK + t];
                const bool a = info.x >= 0, b = info.y >= 0;
                const int nrows = info.z & 0xFF;
                const int ncc = (info.z >> 8) == NB2_ITEM_CONTACTS ? nrows / 3 : 0;
                // the mutable data of the group, all loads in flight together: mj_lambda of both bodies and
                // the group's impulses (same thread wrote them a sweep ago)
                Lam la, lb;
                la.im = lb.im = 0.f;
                if (a) la = load_lam(lam, info.x);
                if (b) lb = load_lam(lam, info.y);
                {
                    float v[12];
#pragma unroll
                    for (int r = 0; r < 12; ++r)
                        if (r < nrows) v[r] = __ldcg(&R.imp[rbase + (unsigned int)r * cnt + g]);
                    refill_prev(e0);
#pragma unroll
                    for (int r = 0; r < 12; ++r)
                        if (r < nrows) simp[r * TPBK + t] = v[r];
                }
                unsigned int cslot = rbase + g;
                // one loop for every group type and both passes: a specialised path per type measured slower
                // (warps mix ground and two-body groups, so the paths serialise)
                if (s == 0) {  // warm start (uniform over the grid): mj_lambda += WJ * cached impulse
#pragma unroll 1
                    for (int r = 0; r < nrows; ++r) {
                        RowPkt k;
                        next_row(r, &k);
                        const float w = k.kind() == NB2_ROW_NONE ? 0.f : k.imp;
                        RowJ J;
                        unpack_pkt(k, true, true, la.im, lb.im, &J);  // unconditional: absent sides are never used
                        if (a) axpy6(w, J.W1, la.v);
                        if (b) axpy6(w, J.W2, lb.v);
                    }
                } else {
#pragma unroll 1
                    for (int r = 0; r < nrows; ++r, cslot += cnt) {
                        RowPkt k;
                        next_row(r, &k);
                        const int kind = k.kind();
                        RowJ J;
                        unpack_pkt(k, true, true, la.im, lb.im, &J);
                        // limits without branches.  Dependent: +-mu * the normal impulse of the previous sweep --
                        // the group's own normal rows (2*ncc..3*ncc-1) come after its friction rows and are not
                        // updated yet; a zero normal impulse clamps to zero, which is what the reference's
                        // special case (sor_prox.rs:251-258) amounts to.
                        float dep = 0.f;
                        if (kind == NB2_ROW_DEPENDENT)
                            dep = r < 2 * ncc ? simp[(2 * ncc + (r >> 1)) * TPBK + t] : __ldcg(&R.imp[k.dep()]);
                        const float lim = k.h.z * dep;
                        const float lo = kind == NB2_ROW_BILATERAL ? k.h.z : (kind == NB2_ROW_DEPENDENT ? -lim : 0.f);
                        const float hi = kind == NB2_ROW_BILATERAL ? k.h.w : (kind == NB2_ROW_DEPENDENT ? lim : NB2_F32_MAX);
                        float d = k.h.x;
                        if (a) d += dot6(J.J1, la.v);
                        if (b) d += dot6(J.J2, lb.v);
                        float ni = fminf(fmaxf(k.imp - k.h.y * d, lo), hi);
                        if (kind == NB2_ROW_NONE) ni = k.imp;
                        const float dl = ni - k.imp;
                        if (a) axpy6(dl, J.W1, la.v);
                        if (b) axpy6(dl, J.W2, lb.v);
                        if (ni != k.imp) __stcg(&R.imp[cslot], ni);
                    }
                }
                if (a) store_lam(lam, info.x, la);
                if (b) store_lam(lam, info.y, lb);
            }
            gb.sync();
#ifdef NB2_TRACE
            if (trace && blockIdx.x == 0 && threadIdx.x == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (s * np + p < 4096) g_trace[s * np + p] = now;
            }
#endif
        }
    }
    cp_async_wait<0>();
#ifdef NB2_TRACE
    if (trace && blockIdx.x == 0 && threadIdx.x == 0)
        for (unsigned int i = 1; i < min((iters + 1) * np, 4096u); ++i)
            if (i / np == (unsigned int)trace)
                printf("sweep %u phase %u groups %u dt_ns %llu\n", i / np, i % np, s_cnt[i % np], g_trace[i] - g_trace[i - 1]);
#else
    (void)trace;
#endif
}

// ------------------------------------------------------------------------------------------
// Bulk-copy row stream (NB2_VELOCITY_KERNEL=3): the same schedule, the same arithmetic, the row ring filled
// by the copy engine instead of by the threads.
//
// The 32 lanes of a warp own 32 consecutive groups of a phase, so row r of those groups is one contiguous
// 512-byte slice in each of the six row planes.  One elected lane arms an mbarrier with the byte count and
// issues six cp.async.bulk (global -> shared, completing on that mbarrier); the warp then waits on the
// mbarrier's phase parity.  Against the per-thread cp.async ring this removes 6 LDGSTS + their address
// arithmetic per thread and row from a kernel that is issue-bound (profiles/r01_notes.md), and the warp
// walks its groups in lockstep (rows beyond a lane's own count are predicated off: groups of a warp hold
// 12 rows each in a pile).  A ring entry is refilled as soon as its values sit in registers.
// g_info travels in registers: the consumer fetches its next group's record before the phase barrier.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int bar,
                                         unsigned long long pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}

struct WarpPos {
    int s;            // sweep, 0 = warm start
    unsigned int p;   // phase
    unsigned int gw;  // first group of the warp's 32-group block within the phase
};
__device__ __forceinline__ bool seek_block(WarpPos& q, unsigned int w0, unsigned int np, int iters, const unsigned int* s_cnt) {
    for (;;) {
        if (q.s > iters) return false;
        if (q.gw < s_cnt[q.p]) return true;
        q.gw = w0;
        if (++q.p == np) {
            q.p = 0;
            ++q.s;
        }
    }
}

#define NB2_BULK_PLANES 6  // 5 jacobian planes + header

template <int D>
__global__ void __launch_bounds__(384, 1) k_velocity_solve_bulk(SchedDev sd, StagedRows R, float4* lam, int iters,
                                                                unsigned int* barrier) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned int TPBK = blockDim.x, NW = TPBK >> 5;
    const unsigned int t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    // [warp][D][6][32] quads | [warp][12][32] impulses | [warp][D] mbarriers | phase tables
    float4* ring_all = reinterpret_cast<float4*>(smem_raw);
    float* simp_all = reinterpret_cast<float*>(ring_all + (size_t)NW * D * NB2_BULK_PLANES * 32);
    unsigned long long* bars_all = reinterpret_cast<unsigned long long*>(simp_all + (size_t)NW * 12 * 32);
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(bars_all + (size_t)NW * D);
    unsigned int* s_rbase = s_cnt + NB2_MAX_COLOURS;
    unsigned int* s_gbase = s_rbase + NB2_MAX_COLOURS;
    const unsigned int np = min(sd.hdr->n_phases, (unsigned int)NB2_MAX_COLOURS);
    if (np == 0) return;
    for (unsigned int i = t; i < np; i += TPBK) {
        s_cnt[i] = sd.ph_count[i];
        s_rbase[i] = sd.ph_rbase[i];
        s_gbase[i] = sd.ph_gbase[i];
    }
    float4* ring = ring_all + (size_t)warp * D * NB2_BULK_PLANES * 32;
    float* simp = simp_all + (size_t)warp * 12 * 32;
    const unsigned int ring_u32 = (unsigned int)__cvta_generic_to_shared(ring);
    const unsigned int bar_u32 = (unsigned int)__cvta_generic_to_shared(bars_all + (size_t)warp * D);
    if (lane == 0) {
#pragma unroll
        for (int e = 0; e < D; ++e) mbar_init(bar_u32 + 8u * e, 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    GridBarrier gb;
    gb.init(barrier);
    const unsigned int w0 = (warp * gridDim.x + blockIdx.x) * 32u;  // block-interleaved warps (interleaved_tid)
    const unsigned int stride = gridDim.x * TPBK;
    const unsigned long long pol = l2_evict_first_policy();
    constexpr unsigned int entry_b = NB2_BULK_PLANES * 32u * 16u, plane_b = 32u * 16u;

    // ---- producer: walks the (sweep, phase, block, row) sequence D entries ahead of the consumer
    WarpPos pc = {0, 0u, w0};
    bool vc = seek_block(pc, w0, np, iters, s_cnt);
    int p_rmax = 0, pr = 0;  // rows of the block being produced, next row
    unsigned int p_slot = 0, p_cnt = 0, p_bytes = 0;
    auto open_block = [&]() {  // all lanes: row count of the block = max over its lanes' groups
        p_cnt = s_cnt[pc.p];
        const unsigned int nvalid = min(32u, p_cnt - pc.gw);
        int nr = 0;
        if (lane < nvalid) nr = __ldg(&sd.g_info[s_gbase[pc.p] + pc.gw + lane]).z & 0xFF;
        p_rmax = __reduce_max_sync(0xffffffffu, nr);
        p_slot = s_rbase[pc.p] + pc.gw;
        p_bytes = nvalid * 16u;
        pr = 0;
    };
    if (vc) open_block();
    int pe = 0;  // ring entry produced next
    auto produce = [&]() {
        while (vc && pr >= p_rmax) {  // next block that has rows
            pc.gw += stride;
            vc = seek_block(pc, w0, np, iters, s_cnt);
            if (vc) open_block();
        }
        if (vc) {
            if (lane == 0) {
                const unsigned int bar = bar_u32 + 8u * (unsigned int)pe;
                const unsigned int dst = ring_u32 + (unsigned int)pe * entry_b;
                mbar_expect_tx(bar, NB2_BULK_PLANES * p_bytes);
#pragma unroll
                for (int k = 0; k < NB2_BULK_PLANES; ++k) bulk_g2s(dst + (unsigned int)k * plane_b, R.q[k] + p_slot, p_bytes, bar, pol);
            }
            p_slot += p_cnt;
            ++pr;
            pe = pe + 1 == D ? 0 : pe + 1;
        }
    };
#pragma unroll 1
    for (int e = 0; e < D; ++e) produce();

    int ce = 0;
    unsigned int cpar = 0;  // parity of the consumer's current pass over the ring
    // next group record of this lane, fetched ahead of the phase barrier
    WarpPos cn = {0, 0u, w0};
    bool cv = seek_block(cn, w0, np, iters, s_cnt);
    int4 info_n = make_int4(-1, -1, 0, 0);
    if (cv && cn.gw + lane < s_cnt[cn.p]) info_n = __ldg(&sd.g_info[s_gbase[cn.p] + cn.gw + lane]);

    for (int s = 0; s <= iters; ++s) {
        for (unsigned int p = 0; p < np; ++p) {
            const unsigned int cnt = s_cnt[p];
            const unsigned int rbase = s_rbase[p];
            for (unsigned int gw = w0; gw < cnt; gw += stride) {
                const unsigned int g = gw + lane;
                const bool active = g < cnt;
                const int4 info = info_n;
                {  // fetch the record of the block after this one
                    cn.gw += stride;
                    cv = seek_block(cn, w0, np, iters, s_cnt);
                    info_n = make_int4(-1, -1, 0, 0);
                    if (cv && cn.gw + lane < s_cnt[cn.p]) info_n = __ldg(&sd.g_info[s_gbase[cn.p] + cn.gw + lane]);
                }
                const bool a = active && info.x >= 0, b = active && info.y >= 0;
                const int nrows = active ? (info.z & 0xFF) : 0;
                const int rmax = __reduce_max_sync(0xffffffffu, nrows);
                const int ncc = (info.z >> 8) == NB2_ITEM_CONTACTS ? nrows / 3 : 0;
                Lam la, lb;
                la.im = lb.im = 0.f;
#pragma unroll
                for (int k = 0; k < 6; ++k) la.v[k] = lb.v[k] = 0.f;
                if (a) la = load_lam(lam, info.x);
                if (b) lb = load_lam(lam, info.y);
                {
                    float v[12];
#pragma unroll
                    for (int r = 0; r < 12; ++r)
                        if (r < nrows) v[r] = __ldcg(&R.imp[rbase + (unsigned int)r * cnt + g]);
#pragma unroll
                    for (int r = 0; r < 12; ++r)
                        if (r < nrows) simp[r * 32 + lane] = v[r];
                }
                unsigned int cslot = rbase + g;
#pragma unroll 1
                for (int r = 0; r < rmax; ++r, cslot += cnt) {
                    mbar_wait(bar_u32 + 8u * (unsigned int)ce, cpar);
                    RowPkt k;
                    const float4* q = ring + (size_t)ce * NB2_BULK_PLANES * 32 + lane;
                    k.q0 = q[0 * 32];
                    k.q1 = q[1 * 32];
                    k.q2 = q[2 * 32];
                    k.q3 = q[3 * 32];
                    k.q4 = q[4 * 32];
                    k.h = q[5 * 32];
                    if (++ce == D) {
                        ce = 0;
                        cpar ^= 1u;
                    }
                    if (r < nrows) {
                        k.imp = simp[r * 32 + lane];
                        const int kind = k.kind();
                        RowJ J;
                        unpack_pkt(k, true, true, la.im, lb.im, &J);
                        if (s == 0) {  // warm start (sor_prox.rs:57-58, 345-435)
                            const float w = kind == NB2_ROW_NONE ? 0.f : k.imp;
                            if (a) axpy6(w, J.W1, la.v);
                            if (b) axpy6(w, J.W2, lb.v);
                        } else {
                            // Dependent limits: +-mu * the normal impulse of the previous sweep (the group's own normal
                            // rows come after its friction rows), exactly as in k_velocity_solve_staged
                            float dep = 0.f;
                            if (kind == NB2_ROW_DEPENDENT)
                                dep = r < 2 * ncc ? simp[(2 * ncc + (r >> 1)) * 32 + lane] : __ldcg(&R.imp[k.dep()]);
                            const float lim = k.h.z * dep;
                            const float lo = kind == NB2_ROW_BILATERAL ? k.h.z : (kind == NB2_ROW_DEPENDENT ? -lim : 0.f);
                            const float hi = kind == NB2_ROW_BILATERAL ? k.h.w : (kind == NB2_ROW_DEPENDENT ? lim : NB2_F32_MAX);
                            float d = k.h.x;
                            if (a) d += dot6(J.J1, la.v);
                            if (b) d += dot6(J.J2, lb.v);
                            float ni = fminf(fmaxf(k.imp - k.h.y * d, lo), hi);
                            if (kind == NB2_ROW_NONE) ni = k.imp;
                            const float dl = ni - k.imp;
                            if (a) axpy6(dl, J.W1, la.v);
                            if (b) axpy6(dl, J.W2, lb.v);
                            if (ni != k.imp) __stcg(&R.imp[cslot], ni);
                        }
                    }
                    // every lane has used its copy of the row: the entry can be refilled by the copy engine
                    __syncwarp();
                    produce();
                }
                if (a) store_lam(lam, info.x, la);
                if (b) store_lam(lam, info.y, lb);
            }
            gb.sync();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Lockstep row stream (NB2_VELOCITY_KERNEL=4): the per-thread cp.async ring of k_velocity_solve_staged with
// the warp-uniform walk of k_velocity_solve_bulk.
//
// In k_velocity_solve_staged every thread walks its own (group, row) sequence, a ring entry per group header
// and per row.  That is fine while all groups of a warp hold the same number of rows; once they do not (a live
// scene: manifolds with 1..4 contacts side by side) a lane with a short group moves on to its next group while
// its neighbours are still in theirs, and from then on the lanes of a warp copy rows of DIFFERENT row indices
// at the same instruction: the 16-byte copies of a warp no longer fall into one 512-byte slice per plane.
// Measured on 4096 x pyramid3 with 3 % of the manifolds ragged: 12.1 -> 16.2 ms.  Here the warp walks in
// lockstep: every lane goes through max-over-the-warp rows per group (short groups commit empty copy groups
// and are predicated off), so lane l always copies row r of group g0 + l: coalesced, whatever the row counts.
// Since the layout hands the slots of a phase out by row count (schedule.cu, k_fill_ginfo) a warp's groups hold equal
// row counts anyway and the free-running rings are faster again (live 100k pile 0.79 against 0.89 ms): this kernel is
// kept as the A/B knob NB2_VELOCITY_KERNEL=4.
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(384, 1) k_velocity_solve_lockstep(SchedDev sd, StagedRows R, float4* lam, int iters,
                                                                    unsigned int* barrier) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned int TPBK = blockDim.x;
    float4* ring = reinterpret_cast<float4*>(smem_raw);                          // [D][6][TPBK]
    float* simp = reinterpret_cast<float*>(ring + D * NB2_BULK_PLANES * TPBK);    // [12][TPBK]
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(simp + 12 * TPBK);
    unsigned int* s_rbase = s_cnt + NB2_MAX_COLOURS;
    unsigned int* s_gbase = s_rbase + NB2_MAX_COLOURS;
    const unsigned int np = min(sd.hdr->n_phases, (unsigned int)NB2_MAX_COLOURS);
    if (np == 0) return;
    const unsigned int t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    for (unsigned int i = t; i < np; i += TPBK) {
        s_cnt[i] = sd.ph_count[i];
        s_rbase[i] = sd.ph_rbase[i];
        s_gbase[i] = sd.ph_gbase[i];
    }
    __syncthreads();
    GridBarrier gb;
    gb.init(barrier);
    const unsigned int w0 = (warp * gridDim.x + blockIdx.x) * 32u;  // block-interleaved warps (interleaved_tid)
    const unsigned int stride = gridDim.x * TPBK;
    const unsigned long long pol = l2_evict_first_policy();
    const unsigned int ring_u32 = (unsigned int)__cvta_generic_to_shared(ring) + t * 16u;
    const unsigned int plane_b = TPBK * 16u, entry_b = NB2_BULK_PLANES * plane_b;

    // ---- producer: D entries (rows) ahead of the consumer, warp-uniform position
    WarpPos pc = {0, 0u, w0};
    bool vc = seek_block(pc, w0, np, iters, s_cnt);
    int p_rmax = 0, p_nr = 0, pr = 0;  // rows of the block being produced (warp max / this lane's), next row
    unsigned int p_slot = 0, p_cnt = 0;
    auto open_block = [&]() {
        p_cnt = s_cnt[pc.p];
        p_nr = 0;
        if (pc.gw + lane < p_cnt) p_nr = __ldg(&sd.g_info[s_gbase[pc.p] + pc.gw + lane]).z & 0xFF;
        p_rmax = __reduce_max_sync(0xffffffffu, p_nr);
        p_slot = s_rbase[pc.p] + pc.gw + lane;
        pr = 0;
    };
    if (vc) open_block();
    int pe = 0;
    auto produce = [&]() {
        while (vc && pr >= p_rmax) {
            pc.gw += stride;
            vc = seek_block(pc, w0, np, iters, s_cnt);
            if (vc) open_block();
        }
        if (vc) {
            if (pr < p_nr) {
                const unsigned int dst = ring_u32 + (unsigned int)pe * entry_b;
                cp_async16_ef(dst + 1u * plane_b, R.q[1] + p_slot, pol);
                cp_async16_ef(dst + 3u * plane_b, R.q[3] + p_slot, pol);
                cp_async16_ef(dst + 4u * plane_b, R.q[4] + p_slot, pol);
                cp_async16_ef(dst + 5u * plane_b, R.q[5] + p_slot, pol);
                cp_async16_ef(dst, R.q[0] + p_slot, pol);
                cp_async16_ef(dst + 2u * plane_b, R.q[2] + p_slot, pol);
            }
            p_slot += p_cnt;
            ++pr;
            pe = pe + 1 == D ? 0 : pe + 1;
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (int e = 0; e < D; ++e) produce();

    int ce = 0;
    WarpPos cn = {0, 0u, w0};
    bool cv = seek_block(cn, w0, np, iters, s_cnt);
    int4 info_n = make_int4(-1, -1, 0, 0);
    if (cv && cn.gw + lane < s_cnt[cn.p]) info_n = __ldg(&sd.g_info[s_gbase[cn.p] + cn.gw + lane]);

    for (int s = 0; s <= iters; ++s) {
        for (unsigned int p = 0; p < np; ++p) {
            const unsigned int cnt = s_cnt[p];
            const unsigned int rbase = s_rbase[p];
            for (unsigned int gw = w0; gw < cnt; gw += stride) {
                const unsigned int g = gw + lane;
                const bool active = g < cnt;
                const int4 info = info_n;
                {
                    cn.gw += stride;
                    cv = seek_block(cn, w0, np, iters, s_cnt);
                    info_n = make_int4(-1, -1, 0, 0);
                    if (cv && cn.gw + lane < s_cnt[cn.p]) info_n = __ldg(&sd.g_info[s_gbase[cn.p] + cn.gw + lane]);
                }
                const bool a = active && info.x >= 0, b = active && info.y >= 0;
                const int nrows = active ? (info.z & 0xFF) : 0;
                const int rmax = __reduce_max_sync(0xffffffffu, nrows);
                const int ncc = (info.z >> 8) == NB2_ITEM_CONTACTS ? nrows / 3 : 0;
                Lam la, lb;
                la.im = lb.im = 0.f;
#pragma unroll
                for (int k = 0; k < 6; ++k) la.v[k] = lb.v[k] = 0.f;
                if (a) la = load_lam(lam, info.x);
                if (b) lb = load_lam(lam, info.y);
                {
                    float v[12];
#pragma unroll
                    for (int r = 0; r < 12; ++r)
                        if (r < nrows) v[r] = __ldcg(&R.imp[rbase + (unsigned int)r * cnt + g]);
#pragma unroll
                    for (int r = 0; r < 12; ++r)
                        if (r < nrows) simp[r * TPBK + t] = v[r];
                }
                unsigned int cslot = rbase + g;
#pragma unroll 1
                for (int r = 0; r < rmax; ++r, cslot += cnt) {
                    cp_async_wait<D - 1>();
                    const float4* q = ring + (size_t)(ce * NB2_BULK_PLANES) * TPBK + t;
                    ce = ce + 1 == D ? 0 : ce + 1;
                    if (r < nrows) {
                        RowPkt k;
                        k.q0 = q[0 * TPBK];
                        k.q1 = q[1 * TPBK];
                        k.q2 = q[2 * TPBK];
                        k.q3 = q[3 * TPBK];
                        k.q4 = q[4 * TPBK];
                        k.h = q[5 * TPBK];
                        k.imp = simp[r * TPBK + t];
                        const int kind = k.kind();
                        RowJ J;
                        unpack_pkt(k, true, true, la.im, lb.im, &J);
                        if (s == 0) {  // warm start (sor_prox.rs:57-58, 345-435)
                            const float w = kind == NB2_ROW_NONE ? 0.f : k.imp;
                            if (a) axpy6(w, J.W1, la.v);
                            if (b) axpy6(w, J.W2, lb.v);
                        } else {
                            float dep = 0.f;
                            if (kind == NB2_ROW_DEPENDENT)
                                dep = r < 2 * ncc ? simp[(2 * ncc + (r >> 1)) * TPBK + t] : __ldcg(&R.imp[k.dep()]);
                            const float lim = k.h.z * dep;
                            const float lo = kind == NB2_ROW_BILATERAL ? k.h.z : (kind == NB2_ROW_DEPENDENT ? -lim : 0.f);
                            const float hi = kind == NB2_ROW_BILATERAL ? k.h.w : (kind == NB2_ROW_DEPENDENT ? lim : NB2_F32_MAX);
                            float d = k.h.x;
                            if (a) d += dot6(J.J1, la.v);
                            if (b) d += dot6(J.J2, lb.v);
                            float ni = fminf(fmaxf(k.imp - k.h.y * d, lo), hi);
                            if (kind == NB2_ROW_NONE) ni = k.imp;
                            const float dl = ni - k.imp;
                            if (a) axpy6(dl, J.W1, la.v);
                            if (b) axpy6(dl, J.W2, lb.v);
                            if (ni != k.imp) __stcg(&R.imp[cslot], ni);
                        }
                    }
                    produce();  // refills the entry just used (its values were consumed above)
                }
                if (a) store_lam(lam, info.x, la);
                if (b) store_lam(lam, info.y, lb);
            }
            gb.sync();
        }
    }
    cp_async_wait<0>();
}

static size_t lockstep_smem(int depth, int tpb) {
    return (size_t)depth * NB2_BULK_PLANES * tpb * 16 + 12 * (size_t)tpb * 4 + 3 * NB2_MAX_COLOURS * 4;
}

int launch_velocity_solve_lockstep(Context* ctx, const SchedDev& sd_in, const Rows& R_in, int tpb, int depth, int blocks) {
    SchedDev sd = sd_in;
    StagedRows R;
    for (int k = 0; k < NB2_ROW_PLANES; ++k) R.q[k] = R_in.jac + (size_t)k * R_in.S;
    R.q[5] = R_in.hdr;
    R.imp = R_in.imp;
    if (depth < 3) depth = 3;
    if (depth > 5) depth = 5;
    while (depth > 3 && lockstep_smem(depth, tpb) > ctx->smem_optin) --depth;
    const size_t smem = lockstep_smem(depth, tpb);
    void* kernel = depth == 3 ? (void*)k_velocity_solve_lockstep<3>
                 : depth == 4 ? (void*)k_velocity_solve_lockstep<4> : (void*)k_velocity_solve_lockstep<5>;
    if (!ctx->lockstep_attr) {
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_lockstep<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_lockstep<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_lockstep<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        ctx->lockstep_attr = true;
    }
    float4* lam = ctx->lam.p;
    int iters = (int)ctx->params.max_velocity_iterations;
    unsigned int* bar = ctx->barrier.p + NB2_BARRIER_VELOCITY;
    void* args[] = {&sd, &R, &lam, &iters, &bar};
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[6], ctx->stream));
    NB2_CUDA(ctx, cudaLaunchCooperativeKernel(kernel, dim3(blocks), dim3(tpb), args, smem, ctx->stream));
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[7], ctx->stream));
    ctx->launches++;
    return NB2_OK;
}

static size_t bulk_smem(int depth, int tpb) {
    const size_t nw = (size_t)tpb / 32;
    return nw * depth * NB2_BULK_PLANES * 32 * 16 + nw * 12 * 32 * 4 + nw * depth * 8 + 3 * NB2_MAX_COLOURS * 4;
}

int launch_velocity_solve_bulk(Context* ctx, const SchedDev& sd_in, const Rows& R_in, int tpb, int depth, int blocks) {
    SchedDev sd = sd_in;
    StagedRows R;
    for (int k = 0; k < NB2_ROW_PLANES; ++k) R.q[k] = R_in.jac + (size_t)k * R_in.S;
    R.q[5] = R_in.hdr;
    R.imp = R_in.imp;
    if (const char* f = getenv("NB2_BULK_DEPTH")) depth = atoi(f);
    if (depth < 3) depth = 3;
    if (depth > 6) depth = 6;
    while (depth > 3 && bulk_smem(depth, tpb) > ctx->smem_optin) --depth;
    const size_t smem = bulk_smem(depth, tpb);
    void* kernel = depth == 3 ? (void*)k_velocity_solve_bulk<3>
                 : depth == 4 ? (void*)k_velocity_solve_bulk<4>
                 : depth == 5 ? (void*)k_velocity_solve_bulk<5> : (void*)k_velocity_solve_bulk<6>;
    if (!ctx->bulk_attr) {
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_bulk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_bulk<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_bulk<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        ctx->bulk_attr = true;
    }
    float4* lam = ctx->lam.p;
    int iters = (int)ctx->params.max_velocity_iterations;
    unsigned int* bar = ctx->barrier.p + NB2_BARRIER_VELOCITY;
    void* args[] = {&sd, &R, &lam, &iters, &bar};
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[6], ctx->stream));
    NB2_CUDA(ctx, cudaLaunchCooperativeKernel(kernel, dim3(blocks), dim3(tpb), args, smem, ctx->stream));
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[7], ctx->stream));
    ctx->launches++;
    return NB2_OK;
}

static size_t staged_smem(int depth, int tpb) {
    return (size_t)depth * NB2_VENTRY * tpb * 16 + 12 * (size_t)tpb * 4 + 3 * NB2_MAX_COLOURS * 4;
}

// Chooses block size and ring depth for the staged kernels: one group per thread and phase while an SM's
// share of a phase fits 384 threads; beyond that 256 threads walk several groups each (measured on
// 4096 x pyramid3, 1.9 M bodies: 128 / 192 / 256 / 320 / 384 threads -> 17.5 / 12.8 / 11.7 / 12.3 / 12.6 ms
// velocity kernel, against 15.0 ms for the register-pipelined kernel).
bool staged_geometry(Context* ctx, int* tpb_out, int* depth_out, int* blocks_out) {
    // groups per phase: the balanced colouring evens the colours out; the colour count is last step's
    // (an asynchronous copy of the schedule header, never waited for), 8 before the first step
    unsigned int np = ctx->host_hdr ? ((volatile SchedHeader*)ctx->host_hdr)->n_phases : 0;
    size_t groups = ctx->host_hdr ? ((volatile SchedHeader*)ctx->host_hdr)->n_groups : 0;
    size_t largest = ctx->host_hdr ? ((volatile SchedHeader*)ctx->host_hdr)->pad[1] : 0;
    if (np == 0 || np > NB2_MAX_COLOURS) np = 8;
    if (groups == 0 || groups > ctx->vs.n_items) groups = ctx->vs.n_items;
    size_t per_phase = (groups + np - 1) / np;
    // an edited (incremental) colouring is not balanced: size the blocks to the LARGEST phase, or its threads
    // would run two groups back to back in every sweep
    if (largest >= per_phase && largest <= groups) per_phase = largest;
    const size_t padded = per_phase + per_phase / 16 + 32;  // balancing tolerance
    // one warp of groups per block until every SM has a block, then wider blocks (a single fat block for
    // small scenes was measured: the grid barrier it saves is worth less than the SMs it gives up)
    size_t blocks = (padded + 31) / 32;
    if (blocks > (size_t)ctx->sm_count) blocks = ctx->sm_count;
    if (blocks < 1) blocks = 1;
    const size_t per_block = (padded + blocks - 1) / blocks;
    int tpb = (int)((per_block + 31) / 32) * 32;
    if (tpb > 384) tpb = 256;
    if (const char* f = getenv("NB2_STAGED_TPB")) tpb = atoi(f);
    if (tpb > 384) tpb = 384;
    if (tpb < 32) tpb = 32;
    // measured on the 100k-box pile (224 threads): depth 3 / 4 / 5 / 7 -> 1.19 / 1.17 / 1.20 / 1.34 ms; a deeper
    // ring fetches the next phase's rows while the current phase still waits for its own
    int depth = NB2_STAGED_DEPTH;
    while (depth > NB2_STAGED_MIN_DEPTH && staged_smem(depth, tpb) > ctx->smem_optin) --depth;
    if (const char* f = getenv("NB2_STAGED_DEPTH")) depth = atoi(f);
    if (depth < NB2_STAGED_MIN_DEPTH) depth = NB2_STAGED_MIN_DEPTH;
    if (depth > NB2_STAGED_MAX_DEPTH) depth = NB2_STAGED_MAX_DEPTH;
    // whatever the overrides asked for must still fit the opt-in shared memory of the device
    while (depth > NB2_STAGED_MIN_DEPTH && staged_smem(depth, tpb) > ctx->smem_optin) --depth;
    while (tpb > 32 && staged_smem(depth, tpb) > ctx->smem_optin) tpb -= 32;
    *blocks_out = (int)blocks;
    *tpb_out = tpb;
    *depth_out = depth;
    return true;
}

int launch_velocity_solve_staged(Context* ctx, const SchedDev& sd_in, const Rows& R_in, int tpb, int depth, int blocks) {
    SchedDev sd = sd_in;
    StagedRows R;
    for (int k = 0; k < NB2_ROW_PLANES; ++k) R.q[k] = R_in.jac + (size_t)k * R_in.S;
    R.q[5] = R_in.hdr;
    R.imp = R_in.imp;
    const size_t smem = staged_smem(depth, tpb);
    void* kernel = depth == 3 ? (void*)k_velocity_solve_staged<3>
                              : (depth == 5 ? (void*)k_velocity_solve_staged<5> : (void*)k_velocity_solve_staged<4>);
    if (!ctx->staged_attr) {
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_staged<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_staged<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_velocity_solve_staged<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        ctx->staged_attr = true;
    }
    float4* lam = ctx->lam.p;
    int iters = (int)ctx->params.max_velocity_iterations;
    unsigned int* bar = ctx->barrier.p + NB2_BARRIER_VELOCITY;
    static const int trace = (getenv("NB2_TRACE_PHASES") ? atoi(getenv("NB2_TRACE_PHASES")) : 0) |
                             (getenv("NB2_NO_EVICT_FIRST") ? 0x40000000 : 0);
    int tr = trace;
    void* args[] = {&sd, &R, &lam, &iters, &bar, &tr};
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[6], ctx->stream));
    NB2_CUDA(ctx, cudaLaunchCooperativeKernel(kernel, dim3(blocks), dim3(tpb), args, smem, ctx->stream));
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[7], ctx->stream));
    ctx->launches++;
    return NB2_OK;
}

// ------------------------------------------------------------------------------------------
// Staged position solve (coloured mode).
//
// Same idea as the staged velocity kernel, at group granularity: everything a contact group needs
// that does not change during the position iterations -- its header (body indices, the two
// collider-to-body poses) and its <= 4 contact records (local points, directions, geometry kinds,
// local normal) -- is copied by cp.async into a per-thread shared-memory entry E groups ahead, across
// phase barriers.  The reference-order kernel chases g_info -> it_src -> chunk_manifold -> manifold ->
// bodies -> contact records, five dependent round trips per group; here only the bodies' state is
// fetched after the barrier, in one round trip.
//
// Entry = 26 quads x TPB lanes: 0..4 group header (p_hdr planes), 5+5*lcc+k contact lcc plane k, 25 g_info.
//
// Early exit: every correction is a pure function of the poses it reads, so a sweep that displaced no body
// would be repeated identically by every later sweep.  Each sweep counts its displacements over the grid
// (one atomic per block that had any); a sweep that ends with the count at zero ends the kernel.  Exact:
// tests/test_gpu_parity.py::test_position_early_exit_is_exact.  (A finer-grained variant -- per-body
// displacement stamps + per-group "clean" stamps, so that settled groups of a partly active pile are skipped
// individually -- was measured on the 100k pile, where 98.7 % of the groups are clean in the first sweep:
// 0.434 ms against 0.397 ms.  The phase time is the latency chain of the few active groups, not the
// throughput of the clean ones; profiles/r02_notes.md.)
// ------------------------------------------------------------------------------------------
#define NB2_PENTRY 26

__global__ void __launch_bounds__(384, 1) k_position_solve_staged(SchedDev sd, PosArrays A, const nb2_joint* __restrict__ joints,
                                                                  const float4* __restrict__ p_hdr, size_t G_stride,
                                                                  const float4* __restrict__ p_row, size_t P_stride,
                                                                  PosParams P, int iters, int E, int rows_div,
                                                                  unsigned int* barrier, int early_exit) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned int TPBK = blockDim.x;
    float4* ring = reinterpret_cast<float4*>(smem_raw);  // [E][NB2_PENTRY][TPBK]
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(ring + (size_t)E * NB2_PENTRY * TPBK);
    unsigned int* s_gbase = s_cnt + NB2_MAX_COLOURS;
    const unsigned int np = min(sd.hdr->n_phases, (unsigned int)NB2_MAX_COLOURS);
    if (np == 0) return;
    for (unsigned int i = threadIdx.x; i < np; i += TPBK) {
        s_cnt[i] = sd.ph_count[i];
        s_gbase[i] = sd.ph_gbase[i];
    }
    __syncthreads();
    GridBarrier gb;
    gb.init(barrier);
    const unsigned int t = threadIdx.x, lane = threadIdx.x & 31u;
    const unsigned int tid = (unsigned int)interleaved_tid();
    const unsigned int stride = gridDim.x * TPBK;
    const unsigned int ring_u32 = (unsigned int)__cvta_generic_to_shared(ring) + t * 16u;
    const unsigned int plane_b = TPBK * 16u, entry_b = NB2_PENTRY * plane_b;
    const int last_it = iters - 1;

    StreamPos pc = {0, 0u, tid}, pn;
    bool vc = seek_group(pc, tid, np, last_it, s_cnt), vn = false;
    int4 ic = make_int4(-1, -1, 0, 0), in_ = ic;
    if (vc) {
        ic = __ldg(&sd.g_info[s_gbase[pc.p] + pc.g]);
        pn = pc;
        pn.g += stride;
        vn = seek_group(pn, tid, np, last_it, s_cnt);
        if (vn) in_ = __ldg(&sd.g_info[s_gbase[pn.p] + pn.g]);
    }
    auto produce = [&](int e) {
        if (vc) {
            const unsigned int dst = ring_u32 + (unsigned int)e * entry_b;
            reinterpret_cast<int4*>(ring)[(e * NB2_PENTRY + 25) * TPBK + t] = ic;
            if ((ic.z >> 8) != NB2_ITEM_JOINT) {
                const unsigned int cnt = s_cnt[pc.p];
                const float4* hsrc = p_hdr + s_gbase[pc.p] + pc.g;
#pragma unroll
                for (int k = 0; k < 5; ++k) cp_async16(dst + (unsigned int)k * plane_b, hsrc + (size_t)k * G_stride);
                const int ncc = (ic.z & 0xFF) / rows_div;
                const float4* rsrc = p_row + (size_t)NB2_CHUNK * s_gbase[pc.p] + pc.g;
                for (int lcc = 0; lcc < ncc; ++lcc, rsrc += cnt) {
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                        cp_async16(dst + (unsigned int)(5 + 5 * lcc + k) * plane_b, rsrc + (size_t)k * P_stride);
                }
            }
            pc = pn;
            ic = in_;
            vc = vn;
            if (vn) {
                pn.g += stride;
                vn = seek_group(pn, tid, np, last_it, s_cnt);
                if (vn) in_ = __ldg(&sd.g_info[s_gbase[pn.p] + pn.g]);
            }
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (int e = 0; e < E; ++e) produce(e);

    int ce = 0;
    unsigned int& s_displaced = s_gbase[NB2_MAX_COLOURS];  // one word behind the phase tables (dynamic shared memory)
    if (threadIdx.x == 0) s_displaced = 0u;
    __syncthreads();
    // [it % 3]: blocks that displaced something in sweep `it`.  Three counters in rotation: the one of sweep it + 2 is
    // cleared after the last barrier of sweep it, a whole sweep (>= 1 barrier) before its first use and after its last
    unsigned int* const sweep_moves = barrier + 2;
    for (int it = 0; it < iters; ++it) {
        unsigned int displaced = 0;
        for (unsigned int p = 0; p < np; ++p) {
            const unsigned int cnt = s_cnt[p];
#ifdef NB2_TRACE
            const bool tracer = blockIdx.x == 0 && threadIdx.x == 0 && it * np + p < 512;
            unsigned long long* tr = g_ptrace + (it * np + p) * 8;
            if (tracer) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[0]));
#endif
            for (unsigned int gw = tid - lane; gw < cnt; gw += stride) {
                if (gw + lane >= cnt) continue;
                cp_async_wait_dyn(E - 1);
#ifdef NB2_TRACE
                if (tracer) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[1]));
#endif
                const int e = ce;
                ce = ce + 1 == E ? 0 : ce + 1;
                const float4* q = ring + (size_t)(e * NB2_PENTRY) * TPBK + t;
                const int4 info = *reinterpret_cast<const int4*>(q + 25 * TPBK);
                PosBody b1, b2;
                if ((info.z >> 8) == NB2_ITEM_JOINT) {
                    const nb2_joint& j = joints[info.w];
                    load_pos_body(A, j.body1, &b1);
                    load_pos_body(A, j.body2, &b2);
                    const Pose o1 = b1.bp.pose, o2 = b2.bp.pose;
                    joint_position(j, &b1, &b2, P);
                    if (b1.dynamic) store_pos_body(A, j.body1, b1);
                    if (b2.dynamic) store_pos_body(A, j.body2, b2);
                    // bitwise comparison: an unchanged pose means no position constraint of the joint fired
                    displaced |= (unsigned int)(o1.t.x != b1.bp.pose.t.x || o1.t.y != b1.bp.pose.t.y || o1.t.z != b1.bp.pose.t.z ||
                                                o1.r.i != b1.bp.pose.r.i || o1.r.j != b1.bp.pose.r.j || o1.r.k != b1.bp.pose.r.k ||
                                                o1.r.w != b1.bp.pose.r.w || o2.t.x != b2.bp.pose.t.x || o2.t.y != b2.bp.pose.t.y ||
                                                o2.t.z != b2.bp.pose.t.z || o2.r.i != b2.bp.pose.r.i || o2.r.j != b2.bp.pose.r.j ||
                                                o2.r.k != b2.bp.pose.r.k || o2.r.w != b2.bp.pose.r.w);
                } else {
                    const float4 h0 = q[0], h1 = q[1 * TPBK], h2 = q[2 * TPBK], h3 = q[3 * TPBK], h4 = q[4 * TPBK];
                    const int body1 = __float_as_int(h0.x), body2 = __float_as_int(h0.y);
                    Pose c1, c2;
                    c1.t = mk3(h1.x, h1.y, h1.z);
                    c1.r = mkq(h1.w, h2.x, h2.y, h2.z);
                    c2.t = mk3(h3.x, h3.y, h3.z);
                    c2.r = mkq(h3.w, h4.x, h4.y, h4.z);
                    const int ncc = (info.z & 0xFF) / rows_div;
                    // colliders sitting at their body's origin (the usual case) need no pose product
                    const bool id1 = h1.x == 0.f && h1.y == 0.f && h1.z == 0.f && h1.w == 0.f && h2.x == 0.f && h2.y == 0.f && h2.z == 1.f;
                    const bool id2 = h3.x == 0.f && h3.y == 0.f && h3.z == 0.f && h3.w == 0.f && h4.x == 0.f && h4.y == 0.f && h4.z == 1.f;
                    load_pos_body(A, body1, &b1);
                    load_pos_body(A, body2, &b2);
#ifdef NB2_TRACE
                    if (tracer) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[2]) : "f"(b1.bp.pose.t.x), "f"(b2.bp.com.z), "f"(b1.inv_i.m[2][2]), "f"(b2.mask[5]));
#endif
                    bool moved1 = false, moved2 = false;
#pragma unroll 1
                    for (int lcc = 0; lcc < ncc; ++lcc) {
                        const float4* rq = q + (size_t)(5 + 5 * lcc) * TPBK;
                        const float4 l1 = rq[0], l2 = rq[1 * TPBK], d1 = rq[2 * TPBK], d2 = rq[3 * TPBK], n1 = rq[4 * TPBK];
                        // update_contact_constraint (nonlinear_sor_prox.rs:156-294)
                        const Pose m1 = id1 ? b1.bp.pose : pose_mul(b1.bp.pose, c1);
                        const Pose m2 = id2 ? b2.bp.pose : pose_mul(b2.bp.pose, c2);
                        ContactEval cev;
                        if (!kinematic_contact(l1, l2, d1, d2, n1, m1, m2, &cev)) continue;
                        const float rhs = clamp_rhs(-cev.depth, false, P);
                        if (rhs >= 0.f) continue;
                        Vec3 w1l = mk3(0.f, 0.f, 0.f), w1a = w1l, w2l = w1l, w2a = w1l;
                        float inv_r = 0.f;
                        pos_fill(b1, cev.world1, false, -cev.normal, &w1l, &w1a, &inv_r);
                        pos_fill(b2, cev.world2, false, cev.normal, &w2l, &w2a, &inv_r);
                        if (inv_r == 0.f) continue;
                        const float impulse = -rhs * (1.f / inv_r);  // solve_unilateral, :137-152
                        if (b1.dynamic) {
                            apply_displacement(&b1.bp, b1.local_com, w1l * impulse, w1a * impulse);
                            moved1 = true;
                        }
                        if (b2.dynamic) {
                            apply_displacement(&b2.bp, b2.local_com, w2l * impulse, w2a * impulse);
                            moved2 = true;
                        }
                    }
#ifdef NB2_TRACE
                    if (tracer) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[3]) : "f"(b1.bp.pose.t.x), "f"(b2.bp.com.z), "r"((int)moved1), "r"((int)moved2));
#endif
                    if (moved1) store_pos_body(A, body1, b1);
                    if (moved2) store_pos_body(A, body2, b2);
                    displaced |= (unsigned int)(moved1 || moved2);
                }
                produce(e);
            }
            if (early_exit && p + 1 == np) {  // last phase of the sweep: publish before the barrier, read after it
                if (displaced) s_displaced = 1u;
                __syncthreads();
                if (threadIdx.x == 0 && s_displaced) {
                    atomicAdd(&sweep_moves[it % 3], 1u);
                    s_displaced = 0u;  // the grid barrier below orders this before the next sweep's writers
                }
            }
#ifdef NB2_TRACE
            gb.sync_traced(tracer ? tr + 4 : nullptr);
#else
            gb.sync();
#endif
        }
        if (early_exit) {
            const unsigned int moves = __ldcg(&sweep_moves[it % 3]);
            if (blockIdx.x == 0 && threadIdx.x == 0) __stcg(&sweep_moves[(it + 2) % 3], 0u);
            if (moves == 0u) break;  // uniform over the grid: nothing moved, nothing will
        }
    }
    cp_async_wait<0>();
#ifdef NB2_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (unsigned int i = 0; i < min((unsigned int)iters * np, 512u); ++i) {
            const unsigned long long* tr = g_ptrace + i * 8;
            printf("pos sweep %u phase %u groups %u | ring %llu bodies %llu contacts %llu block-arrive %llu fence %llu grid %llu | total %llu ns\n",
                   i / np, i % np, s_cnt[i % np], tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4],
                   tr[6] - tr[5], tr[6] - tr[0]);
        }
#endif
}

int launch_position_solve_staged(Context* ctx, const SchedDev& sd_in, const PosArrays& A_in, const PosParams& P_in,
                                 int rows_div, int tpb, int blocks) {
    SchedDev sd = sd_in;
    PosArrays A = A_in;
    PosParams P = P_in;
    // entries per thread: 1 (refilled as soon as its group is done, i.e. while the block waits at the barrier)
    // measured best on the 100k-box pile: 0.52 ms vs 0.54 ms with 2
    int E = 1;
    auto smem_of = [&](int e) { return (size_t)e * NB2_PENTRY * tpb * 16 + 2 * NB2_MAX_COLOURS * 4 + 16; };
    while (E > 1 && smem_of(E) > ctx->smem_optin) --E;
    if (const char* f = getenv("NB2_STAGED_PENTRIES")) E = atoi(f);
    if (smem_of(E) > ctx->smem_optin) return NB2_STAGED_NOT_APPLICABLE;  // caller falls back to the plain kernel
    if (!ctx->staged_pos_attr) {
        NB2_CUDA(ctx, cudaFuncSetAttribute(k_position_solve_staged, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)ctx->smem_optin));
        ctx->staged_pos_attr = true;
    }
    const nb2_joint* joints = ctx->joints.p;
    const float4* phdr = ctx->p_hdr.p;
    size_t gstride = ctx->n_ghdr_max;
    const float4* prow = ctx->p_row.p;
    size_t pstride = ctx->n_pslots_max;
    int iters = (int)ctx->params.max_position_iterations;
    unsigned int* bar = ctx->barrier.p + NB2_BARRIER_POSITION;
    size_t smem = smem_of(E);
    int early = ctx->pos_early_exit ? 1 : 0;
    void* args[] = {&sd, &A, &joints, &phdr, &gstride, &prow, &pstride, &P, &iters, &E, &rows_div, &bar, &early};
    NB2_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_position_solve_staged, dim3(blocks), dim3(tpb), args, smem, ctx->stream));
    ctx->launches++;
    return NB2_OK;
}

template <typename K>
static int coop_limit_c(Context* ctx, K kernel, int* cache) {
    if (*cache > 0) return NB2_OK;
    int per_sm = 0;
    NB2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SOLVE_TPB, 0));
    if (per_sm < 1) return set_error(ctx, NB2_ERR_CUDA, "cooperative kernel does not fit on an SM");
    if (per_sm > 8) per_sm = 8;
    *cache = per_sm * ctx->sm_count;
    return NB2_OK;
}

int launch_velocity_solve_coloured(Context* ctx, const SchedDev& sd_in, const Rows& R_in, const CompactArrays& CA_in) {
    SchedDev sd = sd_in;
    Rows R = R_in;
    CompactArrays CA = CA_in;
    NB2_TRY(coop_limit_c(ctx, k_velocity_solve_coloured, &ctx->coop_blocks_col));
    int iters = (int)ctx->params.max_velocity_iterations;
    unsigned int* bar = ctx->barrier.p + NB2_BARRIER_VELOCITY;
    size_t want = (ctx->vs.n_items + SOLVE_TPB - 1) / SOLVE_TPB;
    int blocks = (int)(want < (size_t)ctx->coop_blocks_col ? want : (size_t)ctx->coop_blocks_col);
    if (blocks < 1) blocks = 1;
    void* args[] = {&sd, &R, &CA, &iters, &bar};
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[6], ctx->stream));
    NB2_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_velocity_solve_coloured, dim3(blocks), dim3(SOLVE_TPB), args, 0,
                                              ctx->stream));
    if (ctx->timers) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[7], ctx->stream));
    ctx->launches++;
    return NB2_OK;
}

}  // namespace nb2
