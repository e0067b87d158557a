// The coloured production velocity kernel.  This translation unit is compiled WITH FMA contraction
// (no -fmad=false): the coloured mode is judged on residual / penetration / energy, not on bit
// equality with the sequential reference order, and half the floating-point instructions matter
// here -- a phase of the 100k-box pile has only ~6 resident warps per SM, so the kernel is bound by
// instruction issue latency, not by memory (profiles/r01_notes.md).
#define NB2_COLOURED_TU 1
#include "solve_compact.cuh"

namespace nb2 {

// Generic rows of one group (joints in coloured mode): plain sequential loads, rare path.
__device__ __noinline__ void generic_group(const Rows& R, float4* lam, int4 info, size_t rbase, unsigned int cnt, size_t g,
                                           bool warm) {
    const bool a = info.x >= 0, b = info.y >= 0;
    const int nrows = info.z & 0xFF;
    Lam la, lb;
    if (a) la = load_lam(lam, info.x);
    if (b) lb = load_lam(lam, info.y);
    for (int r = 0; r < nrows; ++r) {
        const size_t slot = rbase + (size_t)r * cnt + g;
        RowPkt k;
        load_pkt(R, slot, a, b, &k);
        if (k.meta.x == NB2_ROW_NONE) continue;
        RowJ J;
        unpack_pkt(k, a, b, &J);
        if (warm) {
            if (k.imp != 0.f) {
                if (a) axpy6(k.imp, J.W1, la.v);
                if (b) axpy6(k.imp, J.W2, lb.v);
            }
            continue;
        }
        const float dep = k.meta.x == NB2_ROW_DEPENDENT ? __ldcg(&R.imp[k.meta.y]) : 0.f;
        const float ni = solve_row(k.meta.x, k.h, k.imp, dep, J, a, b, &la, &lb);
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
    unsigned int* bar = ctx->barrier.p;
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
