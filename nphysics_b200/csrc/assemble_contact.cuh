// Device helpers of the constraint-row assembly, shared by assemble.cu (joints, the reference-order contact
// kernel, impulse caching; compiled with -fmad=false so that reference-order rows are the oracle's bit for bit) and
// assemble_coloured.cu (the coloured-mode group kernel; FMA contraction allowed, like the coloured solve kernels).
#pragma once
#include "solver.cuh"

namespace nb2 {

static const int ASM_TPB = 128;

// Read-only view of a schedule.
struct SchedView {
    const int* it_phase;
    const int* it_slot;
    const int* it_type;
    const unsigned int* ph_count;
    const unsigned int* ph_gbase;
    const unsigned int* ph_rbase;
    const int4* g_info;
    const int* it_src;
    const SchedHeader* hdr;
    unsigned int max_phases;
    __device__ __forceinline__ unsigned int phase_of(size_t item) const {
        return min((unsigned int)it_phase[item], max_phases - 1);
    }
    __device__ __forceinline__ size_t row_slot(size_t item, int r) const {
        unsigned int p = phase_of(item);
        return (size_t)ph_rbase[p] + (size_t)r * ph_count[p] + (size_t)it_slot[item];
    }
    __device__ __forceinline__ size_t pos_slot(size_t item, int lcc) const {
        unsigned int p = phase_of(item);
        return (size_t)NB2_CHUNK * ph_gbase[p] + (size_t)lcc * ph_count[p] + (size_t)it_slot[item];
    }
};
static inline SchedView view_of(const Sched& s) {
    SchedView v;
    v.it_phase = s.it_phase.p;
    v.it_slot = s.it_slot.p;
    v.it_type = s.it_type.p;
    v.ph_count = s.ph_count.p;
    v.ph_gbase = s.ph_gbase.p;
    v.ph_rbase = s.ph_rbase.p;
    v.g_info = s.g_info.p;
    v.it_src = s.it_src.p;
    v.hdr = s.hdr.p;
    v.max_phases = (unsigned int)s.max_phases;
    return v;
}

struct BodyArrays {
    const nb2_body* raw;
    ConstPoseQuads pos_t;
    ConstPoseQuads pos_q;
    const float4* vel;
    const float4* com_im;
    const float4* inv_i;
    const float4* ext;
};

// What one side of a row needs from its body.
struct BodySide {
    int status;
    Vec3 com;
    float inv_mass;
    Mat3 inv_i;
    float v[6];
    float e[6];
    float mask[6];
};
__device__ __forceinline__ void load_side(const BodyArrays& B, int idx, BodySide* s) {
    // jacobian_mask[6], status, flags are the last two quads of the 176-byte record
    const float4* rq = reinterpret_cast<const float4*>(&B.raw[idx]);
    const float4 m0 = __ldg(rq + 9), m1 = __ldg(rq + 10);
    s->status = __float_as_int(m1.z);
    float4 c = B.com_im[idx];
    s->com = f4_xyz(c);
    s->inv_mass = c.w;
    float4 r0 = B.inv_i[3 * idx], r1 = B.inv_i[3 * idx + 1], r2 = B.inv_i[3 * idx + 2];
    s->inv_i.m[0][0] = r0.x; s->inv_i.m[0][1] = r0.y; s->inv_i.m[0][2] = r0.z;
    s->inv_i.m[1][0] = r1.x; s->inv_i.m[1][1] = r1.y; s->inv_i.m[1][2] = r1.z;
    s->inv_i.m[2][0] = r2.x; s->inv_i.m[2][1] = r2.y; s->inv_i.m[2][2] = r2.z;
    float4 vl = B.vel[2 * idx], va = B.vel[2 * idx + 1];
    s->v[0] = vl.x; s->v[1] = vl.y; s->v[2] = vl.z; s->v[3] = va.x; s->v[4] = va.y; s->v[5] = va.z;
    float4 el = B.ext[2 * idx], ea = B.ext[2 * idx + 1];
    s->e[0] = el.x; s->e[1] = el.y; s->e[2] = el.z; s->e[3] = ea.x; s->e[4] = ea.y; s->e[5] = ea.z;
    s->mask[0] = m0.x; s->mask[1] = m0.y; s->mask[2] = m0.z; s->mask[3] = m0.w; s->mask[4] = m1.x; s->mask[5] = m1.y;
}

__device__ __forceinline__ float dot6_seq(const float* a, const float* b) {
    float res = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) res += a[k] * b[k];
    return res;
}

// RigidBody::fill_constraint_geometry (rigid_body.rs:672-722) for one side.
// J / WJ are left zero for non-dynamic sides.
__device__ __forceinline__ void fill_side(const BodySide& s, Vec3 point, bool angular, Vec3 dir, float* J, float* WJ,
                                          float* inv_r, float* out_vel, bool with_vel) {
    Vec3 pos = point - s.com;
    Vec3 fl = angular ? mk3(0.f, 0.f, 0.f) : dir;
    Vec3 fa = angular ? dir : cross3(pos, dir);
    float f[6] = {fl.x, fl.y, fl.z, fa.x, fa.y, fa.z};
    if (s.status == NB2_BODY_KINEMATIC) {
        if (with_vel) *out_vel += dot6_seq(f, s.v);
    } else if (s.status == NB2_BODY_DYNAMIC) {
        float mf[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) mf[k] = f[k] * s.mask[k];
        Vec3 wl = mk3(mf[0], mf[1], mf[2]) * s.inv_mass;
        Vec3 wa = mat_vec(s.inv_i, mk3(mf[3], mf[4], mf[5]));
#pragma unroll
        for (int k = 0; k < 6; ++k) J[k] = mf[k];
        WJ[0] = wl.x; WJ[1] = wl.y; WJ[2] = wl.z; WJ[3] = wa.x; WJ[4] = wa.y; WJ[5] = wa.z;
        *inv_r += s.inv_mass + dot3(mk3(mf[3], mf[4], mf[5]), wa);
        if (with_vel) {
            *out_vel += dot6_seq(f, s.v);
            *out_vel += dot6_seq(mf, s.e);
        }
    }
}

struct RowOut {
    float4* jac;   // [NB2_ROW_PLANES][n_slots_max], layout in solve_common.cuh
    float4* hdr;
    float* imp;
    size_t n_slots_max;
};
__device__ __forceinline__ void write_row(const RowOut& o, size_t slot, const float* J1, const float* J2,
                                          const float* W1, const float* W2, float rhs, float r, float lo, float hi,
                                          int kind, int dep, float impulse) {
    // streaming stores: half a gigabyte of rows must not evict the bodies, manifolds and hash table
    // the other threads of this kernel are still reading through L2.  The linear part of WJ is not
    // stored (the solve kernels rebuild it as J.lin * inv_mass, the product fill_side formed).
    const size_t S = o.n_slots_max;
    __stcs(&o.jac[0 * S + slot], make_float4(J1[0], J1[1], J1[2], J1[3]));
    __stcs(&o.jac[1 * S + slot], make_float4(J1[4], J1[5], J2[0], J2[1]));
    __stcs(&o.jac[2 * S + slot], make_float4(J2[2], J2[3], J2[4], J2[5]));
    __stcs(&o.jac[3 * S + slot], make_float4(W1[3], W1[4], W1[5], W2[3]));
    __stcs(&o.jac[4 * S + slot], make_float4(W2[4], W2[5], __int_as_float(kind), __int_as_float(dep)));
    __stcs(&o.hdr[slot], make_float4(rhs, r, lo, hi));
    o.imp[slot] = impulse;
}

// helper::constraint_pair_geometry (helper.rs:53-135) + row emission.
__device__ __forceinline__ void emit_pair_row(const RowOut& o, size_t slot, const BodySide& s1, const BodySide& s2,
                                              Vec3 c1, Vec3 c2, bool angular, Vec3 dir, float rhs0, float* rhs_out,
                                              float* r_out, float* J1, float* J2, float* W1, float* W2) {
#pragma unroll
    for (int k = 0; k < 6; ++k) J1[k] = J2[k] = W1[k] = W2[k] = 0.f;
    float inv_r = 0.f;
    float rhs = rhs0;
    fill_side(s1, c1, angular, dir, J1, W1, &inv_r, &rhs, true);
    fill_side(s2, c2, angular, -dir, J2, W2, &inv_r, &rhs, true);
    *r_out = inv_r != 0.f ? 1.f / inv_r : 1.f;
    *rhs_out = rhs;
    (void)o;
    (void)slot;
}

// ---------------------------------------------------------------- impulse cache (hash)
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
// Open-addressing table keyed by the 64-bit contact key; the cached impulses sit in a parallel
// float4 array at the same index, so a lookup issues both loads at once (one round trip per probe).
__device__ __forceinline__ bool ht_lookup(const unsigned long long* __restrict__ keys, const float4* __restrict__ imps,
                                          size_t cap, unsigned long long key, float4* out) {
    if (cap == 0) return false;
    size_t h = (size_t)mix64(key) & (cap - 1);
    for (size_t probe = 0; probe < cap; ++probe) {
        const unsigned long long k = keys[h];
        const float4 v = imps[h];
        if (k == key) {
            *out = v;
            return true;
        }
        if (k == 0ull) return false;
        h = (h + 1) & (cap - 1);
    }
    return false;
}
__device__ __forceinline__ void ht_insert(unsigned long long* keys, float4* imps, size_t cap, unsigned long long key,
                                          float4 val) {
    size_t h = (size_t)mix64(key) & (cap - 1);
    for (size_t probe = 0; probe < cap; ++probe) {
        unsigned long long prev = atomicCAS(&keys[h], 0ull, key);
        if (prev == 0ull || prev == key) {
            imps[h] = val;
            return;
        }
        h = (h + 1) & (cap - 1);
    }
}

// The cache of the previous step as assembly sees it.  Contacts that kept their key AND their index (the
// steady state of a resting scene) are served straight from the per-contact arrays; the hash table over
// the previous step's keys is only built -- on device, by kernels that return at once otherwise -- on
// steps where some contact misses that fast path (k_cache_probe sets *need_hash).
struct ImpulseCacheView {
    const unsigned long long* ckey_prev;  // key of contact i of the previous step
    const float4* imp_prev;               // its impulses (normal, tangent 1, tangent 2)
    unsigned int n_prev;
    const unsigned long long* ht_keys;
    const float4* ht_imps;
    size_t ht_cap;
    unsigned int* need_hash;
    // the device producer's ids (4 p + i + 1) never leave the four slots of their pair: a contact that is in none
    // of them was not there last step, and the hash table need not be built to find that out
    int chunk_local_ids;
};
// fast path: both loads in flight together, one round trip
__device__ __forceinline__ bool cache_fast_path(const ImpulseCacheView& C, unsigned int ci, unsigned long long key, float4* out) {
    if (ci >= C.n_prev) return false;
    const unsigned long long pk = C.ckey_prev[ci];
    const float4 pv = C.imp_prev[ci];
    if (pk != key) return false;
    *out = pv;
    return true;
}
// A contact that is not where it was may still sit in one of the four slots of its own chunk: the device
// producer compacts a pair's kept corners to the front of the pair's slots, so a corner that drops out shifts
// its siblings by one.  Looking there first keeps such steps off the hash path (which clears and rebuilds a
// table over all contacts).
__device__ __forceinline__ bool cache_chunk_path(const ImpulseCacheView& C, unsigned int chunk_first, unsigned int ci,
                                                 unsigned long long key, float4* out) {
#pragma unroll
    for (unsigned int k = 0; k < NB2_CHUNK; ++k) {
        const unsigned int j = chunk_first + k;
        if (j == ci || j >= C.n_prev) continue;
        if (C.ckey_prev[j] == key) {
            *out = C.imp_prev[j];
            return true;
        }
    }
    return false;
}

struct ManifoldConsts {
    float margin1, margin2, friction, restitution;
    Vec3 surf;
};
__device__ __forceinline__ ManifoldConsts manifold_consts(const nb2_manifold& mf) {
    ManifoldConsts K;
    K.margin1 = mf.margin1;
    K.margin2 = mf.margin2;
    K.friction = mf.friction;
    K.restitution = mf.restitution;
    K.surf = mk3(mf.surface_velocity[0], mf.surface_velocity[1], mf.surface_velocity[2]);
    return K;
}
struct ContactSlots {
    size_t n, t1, t2, p;  // row slots of the normal / tangent rows, position slot
};
struct alignas(16) ContactQuads {
    float4 q[7];
};
// the 112-byte contact record as seven quads in one go
__device__ __forceinline__ void load_contact(const nb2_contact* contacts, unsigned int ci, ContactQuads* cq) {
    const float4* cp = reinterpret_cast<const float4*>(&contacts[ci]);
#pragma unroll
    for (int k = 0; k < 7; ++k) cq->q[k] = __ldg(cp + k);
}

// The three velocity rows and the position row of one contact (signorini_coulomb_pyramid_model.rs:56-224,
// signorini_model.rs:37-197), written into the given slots.  q1 = orientation of body 1.
__device__ __forceinline__ void assemble_contact(const RowOut& out, const BodySide& s1, const BodySide& s2,
                                                 const ManifoldConsts& K, const nb2_contact& c, Quat q1, float4 cached,
                                                 const ContactSlots& S, bool compact, float4* c_geo, float4* p_row,
                                                 size_t n_pslots_max, float warmstart_coeff,
                                                 float restitution_threshold, float inv_dt, int model) {
    const size_t slot_n = S.n, slot_t1 = S.t1, slot_t2 = S.t2, pslot = S.p;
    if (model == NB2_CONTACT_SIGNORINI && !(c.depth + K.margin1 + K.margin2 >= 0.f)) {
        // SignoriniModel::is_constraint_active (signorini_model.rs:141-150, applied at :230-232): an inactive
        // contact makes no row at all.  Its slot holds a NONE row whose impulse is the cached one, so that the
        // caching pass carries it to the next step (the model's cache never forgets, :285-297), and a position
        // row no kinematic matches.
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t S_ = out.n_slots_max, P_ = n_pslots_max;
#pragma unroll
        for (int k = 0; k < 4; ++k) out.jac[(size_t)k * S_ + slot_n] = z;
        out.jac[4 * S_ + slot_n] = make_float4(0.f, 0.f, __int_as_float(NB2_ROW_NONE), 0.f);
        out.hdr[slot_n] = z;
        out.imp[slot_n] = cached.x;
#pragma unroll
        for (int k = 0; k < 5; ++k) p_row[(size_t)k * P_ + pslot] = z;
        p_row[2 * P_ + pslot] = make_float4(0.f, 0.f, 0.f, __int_as_float(3));  // geometry tags outside {point, line, plane}
        p_row[3 * P_ + pslot] = make_float4(0.f, 0.f, 0.f, __int_as_float(3));
        return;
    }
    const Vec3 n = mk3(c.normal[0], c.normal[1], c.normal[2]);
    const Vec3 world1 = mk3(c.world1[0], c.world1[1], c.world1[2]);
    const Vec3 world2 = mk3(c.world2[0], c.world2[1], c.world2[2]);
    const Vec3 surf = K.surf;

    // ---- non-penetration row (signorini_model.rs:65-137)
    const Vec3 center1 = world1 + n * K.margin1;
    const Vec3 center2 = world2 - n * K.margin2;
    float J1[6], J2[6], W1[6], W2[6], rhs, r;
    emit_pair_row(out, slot_n, s1, s2, center1, center2, false, -n, dot3(n, surf), &rhs, &r, J1, J2, W1, W2);
    if (rhs <= -restitution_threshold) rhs += K.restitution * rhs;
    float depth = c.depth + K.margin1 + K.margin2;
    if (depth < 0.f) rhs += (-depth) * inv_dt;
    const float rhs_n = rhs, r_n = r;
    if (!compact)
        write_row(out, slot_n, J1, J2, W1, W2, rhs, r, 0.f, NB2_F32_MAX, NB2_ROW_UNILATERAL, 0,
                  cached.x * warmstart_coeff);

    if (model != NB2_CONTACT_SIGNORINI) {
    // ---- friction pyramid rows (signorini_coulomb_pyramid_model.rs:131-216)
    Vec3 t1, t2;
    tangent_basis(n, &t1, &t2);
    emit_pair_row(out, slot_t1, s1, s2, center1, center2, false, t1, dot3(t1, surf), &rhs, &r, J1, J2, W1, W2);
    const float rhs_t1 = rhs, r_t1 = r;
    if (!compact)
        write_row(out, slot_t1, J1, J2, W1, W2, rhs, r, K.friction, 0.f, NB2_ROW_DEPENDENT, (int)slot_n,
                  cached.y * warmstart_coeff);
    emit_pair_row(out, slot_t2, s1, s2, center1, center2, false, t2, dot3(t2, surf), &rhs, &r, J1, J2, W1, W2);
    if (!compact)
        write_row(out, slot_t2, J1, J2, W1, W2, rhs, r, K.friction, 0.f, NB2_ROW_DEPENDENT, (int)slot_n,
                  cached.z * warmstart_coeff);
    if (compact) {
        // Compact record: the solve kernel rebuilds J = mask*(d, p x d) and WJ = M^-1 J with the
        // very expressions of fill_side, so the rows it iterates are bit-identical to the
        // 132-byte rows of the reference-order layout at a fifth of the bytes.
        const Vec3 p1 = center1 - s1.com, p2 = center2 - s2.com;
        const size_t P_ = n_pslots_max;
        c_geo[0 * P_ + pslot] = make_float4(p1.x, p1.y, p1.z, rhs_n);
        c_geo[1 * P_ + pslot] = make_float4(p2.x, p2.y, p2.z, rhs_t1);
        c_geo[2 * P_ + pslot] = make_float4(n.x, n.y, n.z, rhs);
        c_geo[3 * P_ + pslot] = make_float4(r_n, r_t1, r, K.friction);
        c_geo[4 * P_ + pslot] = make_float4(cached.x * warmstart_coeff, cached.y * warmstart_coeff,
                                            cached.z * warmstart_coeff, 1.f);
    }
    }

    // ---- position row (signorini_model.rs:153-197)
    const Vec3 normal1 = quat_inv_rotate(q1, n);
    const size_t P = n_pslots_max;
    __stcs(&p_row[0 * P + pslot], make_float4(c.local1[0], c.local1[1], c.local1[2], c.dilation1 + K.margin1));
    __stcs(&p_row[1 * P + pslot], make_float4(c.local2[0], c.local2[1], c.local2[2], c.dilation2 + K.margin2));
    __stcs(&p_row[2 * P + pslot], make_float4(c.dir1[0], c.dir1[1], c.dir1[2], __int_as_float((int)c.geom1)));
    __stcs(&p_row[3 * P + pslot], make_float4(c.dir2[0], c.dir2[1], c.dir2[2], __int_as_float((int)c.geom2)));
    __stcs(&p_row[4 * P + pslot], make_float4(normal1.x, normal1.y, normal1.z, 0.f));
}

// Coloured mode: one thread per contact GROUP (the <= 4 contacts of a manifold chunk), enumerated in ELL
// order.  The thread walks g_info -> chunk -> manifold once, loads the two bodies once for all its
// contacts (a per-contact thread re-reads them four times; at 1.9 M bodies they no longer sit in L2)
// and writes rows that are consecutive over the group index: consecutive threads hit consecutive
// 16-byte words of every plane.  It also writes the group header of the staged position kernel.
// 2 blocks of 128 per SM (254 registers, no spill) measured best: assembly stage 0.346 ms on the 100k pile and
// 4.8 ms at 1.9 M bodies, against 0.374 / 5.1 at 3 blocks (168 registers, 212 bytes of spill) and 0.411 / 5.8 at 4

}  // namespace nb2
