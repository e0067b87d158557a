// Coloured-mode contact assembly: one thread per contact group, rows written straight into their ELL slots.
// Compiled with -fmad=false like assemble.cu.  Measured (profiles/r02_notes.md): contracting this translation unit
// to FMA leaves the stage where it was (0.347 ms against 0.343 ms on the 100k pile: the kernel waits on its
// scattered loads, not on its arithmetic) and moves the trajectory of the buckling 50x200 wall outside its stated
// tolerance, so the reference's multiply-add order is kept here as well.
#include "assemble_contact.cuh"

namespace nb2 {

static const int TPB = ASM_TPB;

#define NB2_ASMG_MINBLOCKS 2
__global__ void __launch_bounds__(TPB, NB2_ASMG_MINBLOCKS) k_assemble_groups(
    unsigned int nC, const nb2_manifold* __restrict__ manifolds, const nb2_contact* __restrict__ contacts,
    const unsigned int* __restrict__ chunk_base, const unsigned int* __restrict__ chunk_manifold, BodyArrays B,
    SchedView vs, RowOut out, float4* p_row, size_t n_pslots_max, float4* p_hdr, size_t n_ghdr_max, float4* c_geo,
    ImpulseCacheView cache, float warmstart_coeff, float restitution_threshold, float inv_dt, int compact_layout,
    int model) {
    const size_t T = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int np = vs.hdr->n_phases;
    if (np == 0 || T >= (size_t)vs.ph_gbase[np]) return;
    unsigned int lo = 0, hi = np;  // largest p with gbase[p] <= T
    while (hi - lo > 1) {
        unsigned int mid = (lo + hi) >> 1;
        if ((size_t)vs.ph_gbase[mid] <= T) lo = mid; else hi = mid;
    }
    const unsigned int p = lo, cnt = vs.ph_count[p];
    const unsigned int g = (unsigned int)(T - vs.ph_gbase[p]);
    const int4 gi = vs.g_info[T];
    if ((gi.z >> 8) != NB2_ITEM_CONTACTS) return;
    const unsigned int chunk = (unsigned int)vs.it_src[gi.w];
    const unsigned int m = chunk_manifold[chunk];
    const unsigned int lchunk = chunk - chunk_base[m];
    const nb2_manifold& mf = manifolds[m];
    const int ncc = min(NB2_CHUNK, (int)mf.num_contacts - (int)(NB2_CHUNK * lchunk));
    const unsigned int ci0 = mf.first_contact + NB2_CHUNK * lchunk;
    const int body1 = mf.body1, body2 = mf.body2;
    {
        const float* k1 = mf.coll1_wrt_body;
        const float* k2 = mf.coll2_wrt_body;
        p_hdr[0 * n_ghdr_max + T] = make_float4(__int_as_float(body1), __int_as_float(body2), __int_as_float((int)m), 0.f);
        p_hdr[1 * n_ghdr_max + T] = make_float4(k1[0], k1[1], k1[2], k1[3]);
        p_hdr[2 * n_ghdr_max + T] = make_float4(k1[4], k1[5], k1[6], 0.f);
        p_hdr[3 * n_ghdr_max + T] = make_float4(k2[0], k2[1], k2[2], k2[3]);
        p_hdr[4 * n_ghdr_max + T] = make_float4(k2[4], k2[5], k2[6], 0.f);
    }
    const ManifoldConsts K = manifold_consts(mf);
    const bool compact = compact_layout != 0;
    const size_t pbase = (size_t)NB2_CHUNK * vs.ph_gbase[p] + g, rb = (size_t)vs.ph_rbase[p] + g;
    if (compact)  // lanes beyond the chunk's contacts: flag their compact records invalid
        for (int lcc = max(ncc, 0); lcc < NB2_CHUNK; ++lcc)
            c_geo[4 * n_pslots_max + pbase + (size_t)lcc * cnt] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ncc <= 0 || ci0 + (unsigned int)ncc > nC) return;
    BodySide s1, s2;
    load_side(B, body1, &s1);
    load_side(B, body2, &s2);
    if (s1.status != NB2_BODY_DYNAMIC && s2.status != NB2_BODY_DYNAMIC) return;
    const Quat q1 = f4_quat(B.pos_q[body1]);
    ContactQuads cur, nxt;
    load_contact(contacts, ci0, &cur);
#pragma unroll 1
    for (int lcc = 0; lcc < ncc; ++lcc) {
        const unsigned int ci = ci0 + (unsigned int)lcc;
        if (lcc + 1 < ncc) load_contact(contacts, ci + 1, &nxt);  // next record in flight while this one is assembled
        const nb2_contact& c = *reinterpret_cast<const nb2_contact*>(&cur);
        float4 cached = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c.key != 0ull) {  // impulse cache (signorini_coulomb_pyramid_model.rs:104-108)
            float4 prev;
            if (cache_fast_path(cache, ci, c.key, &prev) || cache_chunk_path(cache, ci0, ci, c.key, &prev)) cached = prev;
            else if (!cache.chunk_local_ids) *cache.need_hash = 1u;  // k_warm_fixup patches this contact's warm start once the table exists
        }
        ContactSlots S;
        S.p = pbase + (size_t)lcc * cnt;
        S.t1 = rb + (size_t)(2 * lcc) * cnt;
        S.t2 = rb + (size_t)(2 * lcc + 1) * cnt;
        S.n = model == NB2_CONTACT_SIGNORINI ? rb + (size_t)lcc * cnt : rb + (size_t)(2 * ncc + lcc) * cnt;
        assemble_contact(out, s1, s2, K, c, q1, cached, S, compact, c_geo, p_row, n_pslots_max, warmstart_coeff,
                         restitution_threshold, inv_dt, model);
        cur = nxt;
    }
}

// Reference order: one thread per contact.

int launch_assemble_groups(Context* ctx, size_t n_items, const BodyArrays& B, const SchedView& vs, const RowOut& out,
                           const ImpulseCacheView& cache) {
    k_assemble_groups<<<(unsigned int)((n_items + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(
        ctx->n_contacts, ctx->manifolds.p, ctx->contacts.p, ctx->chunk_base.p, ctx->chunk_manifold.p, B, vs, out, ctx->p_row.p,
        ctx->n_pslots_max, ctx->p_hdr.p, ctx->n_ghdr_max, ctx->c_geo.p, cache, ctx->params.warmstart_coeff,
        ctx->params.restitution_velocity_threshold, ctx->inv_dt, ctx->step_layout, ctx->contact_model);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

}  // namespace nb2
