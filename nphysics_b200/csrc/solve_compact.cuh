// Compact contact-group device code (coloured mode), shared by solve.cu (diagnostics, built with
// -fmad=false) and solve_coloured.cu (the production kernel, built with FMA contraction).
#pragma once
#include "solve_common.cuh"

namespace nb2 {

// ------------------------------------------------------------------------------------------
// Compact contact groups (coloured mode).  A contact travels as 5 float4 (80 B):
//   (p1, rhs_n) (p2, rhs_t1) (n, rhs_t2) (r_n, r_t1, r_t2, mu) (imp_n, imp_t1, imp_t2, valid)
// and its three rows are rebuilt in registers exactly as RigidBody::fill_constraint_geometry builds
// them (src/object/rigid_body.rs:672-722): J = mask * (d, p x d), WJ = M^-1 J.  All 20 loads of a
// group (<= 4 contacts) and the 12 body loads are issued before the first use, so one thread has
// its whole working set in flight at once: a phase owns only ~200 groups per SM and latency is
// hidden by loads in flight per thread, not by occupancy.
// ------------------------------------------------------------------------------------------
struct CompactBody {
    bool dyn;
    int idx;
    float inv_mass;
    Mat3 inv_i;
    Lam l;
};
struct CompactArrays {
    const nb2_body* raw;
    const float4* com_im;
    const float4* inv_i;
    float4* lam;
    float4* c_geo;
    size_t P;      // plane stride (= n_pslots_max)
    int any_mask;
};
__device__ __forceinline__ void load_compact_body(const CompactArrays& A, int idx, CompactBody* o) {
    o->dyn = idx >= 0;
    o->idx = idx;
    if (!o->dyn) return;
    const float4 c = __ldg(&A.com_im[idx]);
    const float4 r0 = __ldg(&A.inv_i[3 * idx]), r1 = __ldg(&A.inv_i[3 * idx + 1]), r2 = __ldg(&A.inv_i[3 * idx + 2]);
    o->l = load_lam(A.lam, idx);
    o->inv_mass = c.w;
    o->inv_i.m[0][0] = r0.x; o->inv_i.m[0][1] = r0.y; o->inv_i.m[0][2] = r0.z;
    o->inv_i.m[1][0] = r1.x; o->inv_i.m[1][1] = r1.y; o->inv_i.m[1][2] = r1.z;
    o->inv_i.m[2][0] = r2.x; o->inv_i.m[2][1] = r2.y; o->inv_i.m[2][2] = r2.z;
}
// one side of a row: J and WJ of body `b` for a unit force `dir` applied at `pos` (relative to the com)
__device__ __forceinline__ void compact_side(const CompactArrays& A, const CompactBody& b, Vec3 pos, Vec3 dir, float* J,
                                             float* W) {
    const Vec3 fa = cross3(pos, dir);
    J[0] = dir.x; J[1] = dir.y; J[2] = dir.z; J[3] = fa.x; J[4] = fa.y; J[5] = fa.z;
    if (A.any_mask) {  // rare: some dof of some body is kinematic (rigid_body.rs:105-121)
        const float* m = A.raw[b.idx].jacobian_mask;
#pragma unroll
        for (int k = 0; k < 6; ++k) J[k] = J[k] * m[k];
    }
    const Vec3 wl = mk3(J[0], J[1], J[2]) * b.inv_mass;
    const Vec3 wa = mat_vec(b.inv_i, mk3(J[3], J[4], J[5]));
    W[0] = wl.x; W[1] = wl.y; W[2] = wl.z; W[3] = wa.x; W[4] = wa.y; W[5] = wa.z;
}
// what: 0 = solve, 1 = warm start, 2 = residual only (returns |prox step|)
__device__ __forceinline__ float compact_row(const CompactArrays& A, int what, CompactBody* b1, CompactBody* b2, Vec3 p1,
                                             Vec3 p2, Vec3 dir, float rhs, float r, int kind, float mu, float dep,
                                             float* impulse) {
    RowJ J;
    if (b1->dyn) compact_side(A, *b1, p1, dir, J.J1, J.W1);
    if (b2->dyn) compact_side(A, *b2, p2, -dir, J.J2, J.W2);
    if (what == 1) {
        if (*impulse != 0.f) {
            if (b1->dyn) axpy6(*impulse, J.W1, b1->l.v);
            if (b2->dyn) axpy6(*impulse, J.W2, b2->l.v);
        }
        return 0.f;
    }
    if (what == 2) {
        float d;
        if (b1->dyn && b2->dyn) d = dot6(J.J1, b1->l.v) + dot6(J.J2, b2->l.v) + rhs;
        else if (b1->dyn) d = dot6(J.J1, b1->l.v) + rhs;
        else d = dot6(J.J2, b2->l.v) + rhs;
        float ni;
        if (kind == NB2_ROW_UNILATERAL) ni = fmaxf(*impulse - r * d, 0.f);
        else ni = clampf(*impulse - r * d, -(mu * dep), mu * dep);
        return fabsf(ni - *impulse);
    }
    *impulse = solve_row(kind, make_float4(rhs, r, mu, 0.f), *impulse, dep, J, b1->dyn, b2->dyn, &b1->l, &b2->l);
    return 0.f;
}
__device__ __forceinline__ void compact_group(const CompactArrays& A, int what, int a, int b, size_t pbase,
                                              unsigned int cnt, size_t g, int ncc, float* res_max, double* res_sq,
                                              unsigned int* res_n) {
    float4 q0[NB2_CHUNK], q1[NB2_CHUNK], q2[NB2_CHUNK], q3[NB2_CHUNK], q4[NB2_CHUNK];
#pragma unroll
    for (int k = 0; k < NB2_CHUNK; ++k) {
        q4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < ncc) {
            const size_t ps = pbase + (size_t)k * cnt + g;
            q4[k] = ldcg4(&A.c_geo[4 * A.P + ps]);
            q0[k] = __ldg(&A.c_geo[0 * A.P + ps]);
            q1[k] = __ldg(&A.c_geo[1 * A.P + ps]);
            q2[k] = __ldg(&A.c_geo[2 * A.P + ps]);
            q3[k] = __ldg(&A.c_geo[3 * A.P + ps]);
        }
    }
    CompactBody b1, b2;
    load_compact_body(A, a, &b1);
    load_compact_body(A, b, &b2);
    // friction rows first: they are limited by the normal impulse of the previous sweep
    // (sor_prox.rs:167-178), i.e. the value q4.x holds on entry
#pragma unroll
    for (int k = 0; k < NB2_CHUNK; ++k) {
        if (k < ncc && q4[k].w != 0.f) {
            const Vec3 n = f4_xyz(q2[k]);
            Vec3 t1, t2;
            tangent_basis(n, &t1, &t2);
            const Vec3 p1 = f4_xyz(q0[k]), p2 = f4_xyz(q1[k]);
            const float r1 = compact_row(A, what, &b1, &b2, p1, p2, t1, q1[k].w, q3[k].y, NB2_ROW_DEPENDENT, q3[k].w,
                                         q4[k].x, &q4[k].y);
            const float r2 = compact_row(A, what, &b1, &b2, p1, p2, t2, q2[k].w, q3[k].z, NB2_ROW_DEPENDENT, q3[k].w,
                                         q4[k].x, &q4[k].z);
            if (what == 2) {
                *res_max = fmaxf(*res_max, fmaxf(r1, r2));
                *res_sq += (double)r1 * r1 + (double)r2 * r2;
                *res_n += 2;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NB2_CHUNK; ++k) {
        if (k < ncc && q4[k].w != 0.f) {
            const float r0 = compact_row(A, what, &b1, &b2, f4_xyz(q0[k]), f4_xyz(q1[k]), -f4_xyz(q2[k]), q0[k].w,
                                         q3[k].x, NB2_ROW_UNILATERAL, 0.f, 0.f, &q4[k].x);
            if (what == 2) {
                *res_max = fmaxf(*res_max, r0);
                *res_sq += (double)r0 * r0;
                *res_n += 1;
            }
        }
    }
    if (what == 2) return;
    if (what == 0) {
#pragma unroll
        for (int k = 0; k < NB2_CHUNK; ++k)
            if (k < ncc && q4[k].w != 0.f) stcg4(&A.c_geo[4 * A.P + pbase + (size_t)k * cnt + g], q4[k]);
    }
    if (b1.dyn) store_lam(A.lam, a, b1.l);
    if (b2.dyn) store_lam(A.lam, b, b2.l);
}


}  // namespace nb2
