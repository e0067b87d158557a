// Shared device helpers of the solve kernels (included by solve.cu and solve_coloured.cu).
#pragma once
#include "solver.cuh"

namespace nb2 {

static const int SOLVE_TPB = 128;

struct SchedDev {
    const unsigned int* ph_count;
    const unsigned int* ph_gbase;
    const unsigned int* ph_rbase;
    const int4* g_info;
    const SchedHeader* hdr;
    const int* it_type;
    const int* it_src;
    const int* it_a;
    const unsigned long long* it_key;
    const int* it_phase;
    const int* it_slot;
    unsigned int max_phases;
};
static inline SchedDev sched_dev(const Sched& s) {
    SchedDev d;
    d.ph_count = s.ph_count.p;
    d.ph_gbase = s.ph_gbase.p;
    d.ph_rbase = s.ph_rbase.p;
    d.g_info = s.g_info.p;
    d.hdr = s.hdr.p;
    d.it_type = s.it_type.p;
    d.it_src = s.it_src.p;
    d.it_a = s.it_a.p;
    d.it_key = s.it_key.p;
    d.it_phase = s.it_phase.p;
    d.it_slot = s.it_slot.p;
    d.max_phases = (unsigned int)s.max_phases;
    return d;
}

// Velocity rows, 100 bytes each: five jacobian quads, a header quad, the impulse.
//   jac[0] = J1[0..3]   jac[1] = J1[4..5] J2[0..1]   jac[2] = J2[2..5]
//   jac[3] = WJ1.ang.xyz, WJ2.ang.x    jac[4] = WJ2.ang.yz, kind (int bits), dependency slot (int bits)
//   hdr    = rhs, r, lo | mu, hi
// The linear half of WJ = M^-1 J is not stored: for a rigid body it is J.lin * inv_mass (exactly the
// product fill_constraint_geometry forms, rigid_body.rs:703-706), and inv_mass rides in the spare
// lane of the body's mj_lambda quad.
struct Rows {
    const float4* jac;
    const float4* hdr;
    float* imp;
    size_t S;  // plane stride
};

struct Lam {
    float v[6];
    float im;  // inverse mass of the body (constant; carried through load/store)
};
__device__ __forceinline__ Lam load_lam(const float4* lam, int b) {
    Lam l;
    float4 a = ldcg4(&lam[2 * b]), c = ldcg4(&lam[2 * b + 1]);
    l.v[0] = a.x; l.v[1] = a.y; l.v[2] = a.z; l.v[3] = c.x; l.v[4] = c.y; l.v[5] = c.z;
    l.im = a.w;
    return l;
}
__device__ __forceinline__ void store_lam(float4* lam, int b, const Lam& l) {
    stcg4(&lam[2 * b], make_float4(l.v[0], l.v[1], l.v[2], l.im));
    stcg4(&lam[2 * b + 1], make_float4(l.v[3], l.v[4], l.v[5], 0.f));
}
__device__ __forceinline__ float dot6(const float* a, const float* b) {
#ifdef NB2_COLOURED_TU
    // production kernel: two independent 3-term FMA chains (half the dependent depth)
    float s0 = a[0] * b[0], s1 = a[3] * b[3];
    s0 = fmaf(a[1], b[1], s0);
    s1 = fmaf(a[4], b[4], s1);
    s0 = fmaf(a[2], b[2], s0);
    s1 = fmaf(a[5], b[5], s1);
    return s0 + s1;
#else
    // nalgebra's order: sequential res += a[k] * b[k]
    float res = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) res += a[k] * b[k];
    return res;
#endif
}
__device__ __forceinline__ void axpy6(float a, const float* x, float* y) {
#pragma unroll
    for (int k = 0; k < 6; ++k) y[k] = a * x[k] + y[k];
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v > lo ? (v < hi ? v : hi) : lo; }

struct RowJ {
    float J1[6], J2[6], W1[6], W2[6];
};
__device__ __forceinline__ int row_kind(const Rows& R, size_t slot) {
    return __float_as_int(__ldg(&R.jac[4 * R.S + slot]).z);
}
// ima / imb: inverse masses of the two bodies (only the present sides are filled)
__device__ __forceinline__ void load_row_j(const Rows& R, size_t slot, bool a, bool b, float ima, float imb, RowJ* o) {
    const float4 q1 = __ldg(&R.jac[1 * R.S + slot]);
    const float4 q3 = __ldg(&R.jac[3 * R.S + slot]);
    if (a) {
        const float4 q0 = __ldg(&R.jac[0 * R.S + slot]);
        o->J1[0] = q0.x; o->J1[1] = q0.y; o->J1[2] = q0.z; o->J1[3] = q0.w; o->J1[4] = q1.x; o->J1[5] = q1.y;
        o->W1[0] = q0.x * ima; o->W1[1] = q0.y * ima; o->W1[2] = q0.z * ima;
        o->W1[3] = q3.x; o->W1[4] = q3.y; o->W1[5] = q3.z;
    }
    if (b) {
        const float4 q2 = __ldg(&R.jac[2 * R.S + slot]);
        const float4 q4 = __ldg(&R.jac[4 * R.S + slot]);
        o->J2[0] = q1.z; o->J2[1] = q1.w; o->J2[2] = q2.x; o->J2[3] = q2.y; o->J2[4] = q2.z; o->J2[5] = q2.w;
        o->W2[0] = q1.z * imb; o->W2[1] = q1.w * imb; o->W2[2] = q2.x * imb;
        o->W2[3] = q3.w; o->W2[4] = q4.x; o->W2[5] = q4.y;
    }
}

// One SORProx row update (sor_prox.rs:181-343) on register-resident mj_lambda.
// Returns the new impulse.
__device__ __forceinline__ float solve_row(int kind, float4 h, float impulse, float dep_impulse, const RowJ& J,
                                           bool a, bool b, Lam* la, Lam* lb) {
    float lo, hi;
    if (kind == NB2_ROW_UNILATERAL) {
        lo = 0.f;
        hi = NB2_F32_MAX;
    } else if (kind == NB2_ROW_BILATERAL) {
        lo = h.z;
        hi = h.w;
    } else {  // Dependent: sor_prox.rs:251-272
        if (dep_impulse == 0.f) {
            if (impulse != 0.f) {
                if (a) axpy6(-impulse, J.W1, la->v);
                if (b) axpy6(-impulse, J.W2, lb->v);
            }
            return 0.f;
        }
        hi = h.z * dep_impulse;
        lo = -hi;
    }
    float d;
    if (a && b)
        d = dot6(J.J1, la->v) + dot6(J.J2, lb->v) + h.x;
    else if (a)
        d = dot6(J.J1, la->v) + h.x;
    else
        d = dot6(J.J2, lb->v) + h.x;
    float ni;
    if (kind == NB2_ROW_UNILATERAL)
        ni = fmaxf(impulse - h.y * d, 0.f);
    else
        ni = clampf(impulse - h.y * d, lo, hi);
    float dl = ni - impulse;
    if (a) axpy6(dl, J.W1, la->v);
    if (b) axpy6(dl, J.W2, lb->v);
    return ni;
}

// A row as it travels from the ELL planes to the update: 5 jacobian quads (kind and dependency slot
// in the last one), header, the row's impulse.
struct RowPkt {
    float4 q0, q1, q2, q3, q4, h;
    float imp;
    __device__ __forceinline__ int kind() const { return __float_as_int(q4.z); }
    __device__ __forceinline__ int dep() const { return __float_as_int(q4.w); }
};
// All loads of a row are issued together and one row ahead of its use (software pipelining):
// per phase an SM owns only ~200 groups, so latency is hidden by loads in flight per thread,
// not by occupancy.  Read-only planes go through ld.global.nc; impulses bypass L1 (ld.cg)
// because other SMs wrote them in an earlier phase.
__device__ __forceinline__ void load_pkt(const Rows& R, size_t slot, bool a, bool b, RowPkt* o) {
    o->h = __ldg(&R.hdr[slot]);
    o->q1 = __ldg(&R.jac[1 * R.S + slot]);
    o->q3 = __ldg(&R.jac[3 * R.S + slot]);
    o->q4 = __ldg(&R.jac[4 * R.S + slot]);
    if (a) o->q0 = __ldg(&R.jac[0 * R.S + slot]);
    if (b) o->q2 = __ldg(&R.jac[2 * R.S + slot]);
    o->imp = __ldcg(&R.imp[slot]);
}
__device__ __forceinline__ void unpack_pkt(const RowPkt& k, bool a, bool b, float ima, float imb, RowJ* o) {
    if (a) {
        o->J1[0] = k.q0.x; o->J1[1] = k.q0.y; o->J1[2] = k.q0.z; o->J1[3] = k.q0.w; o->J1[4] = k.q1.x; o->J1[5] = k.q1.y;
        o->W1[0] = k.q0.x * ima; o->W1[1] = k.q0.y * ima; o->W1[2] = k.q0.z * ima;
        o->W1[3] = k.q3.x; o->W1[4] = k.q3.y; o->W1[5] = k.q3.z;
    }
    if (b) {
        o->J2[0] = k.q1.z; o->J2[1] = k.q1.w; o->J2[2] = k.q2.x; o->J2[3] = k.q2.y; o->J2[4] = k.q2.z; o->J2[5] = k.q2.w;
        o->W2[0] = k.q1.z * imb; o->W2[1] = k.q1.w * imb; o->W2[2] = k.q2.x * imb;
        o->W2[3] = k.q3.w; o->W2[4] = k.q4.x; o->W2[5] = k.q4.y;
    }
}
// Warps are dealt to groups block-interleaved (warp w of block b is global warp w*gridDim+b) so a
// phase with fewer groups than threads still spreads evenly over all SMs.
__device__ __forceinline__ size_t interleaved_tid() {
    return ((size_t)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
}


}  // namespace nb2
