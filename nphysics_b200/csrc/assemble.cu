// Constraint-row assembly on device and impulse caching.
//
// assemble_contact_rows replaces SignoriniCoulombPyramidModel::constraints
// (src/solver/signorini_coulomb_pyramid_model.rs:56-224) with SignoriniModel::
// build_velocity_constraint / build_position_constraint (signorini_model.rs:37-197),
// helper::constraint_pair_geometry (helper.rs:53-135) and RigidBody::
// fill_constraint_geometry (src/object/rigid_body.rs:672-722) inlined: one thread per
// contact emits the non-penetration row, the two friction-pyramid rows and the position row
// straight into the ELL slots chosen by the schedule.
// assemble_joint_rows replaces each *_constraint.rs::velocity_constraints (one thread per joint).
// cache_impulses replaces the two cache_impulses passes (signorini_coulomb_pyramid_model.rs:
// 226-261, ball_constraint.rs:132-144 and friends).
#include "solver.cuh"

namespace nb2 {

static const int TPB = 128;
static inline unsigned int nblk(size_t n) { return (unsigned int)((n + TPB - 1) / TPB); }

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
static SchedView view_of(const Sched& s) {
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
    const float4* pos_t;
    const float4* pos_q;
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
__global__ void k_hash_clear(const unsigned int* __restrict__ need_hash, unsigned long long* keys, size_t cap) {
    if (*need_hash == 0u) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) keys[i] = 0ull;
}
__global__ void k_hash_build(const unsigned int* __restrict__ need_hash, const unsigned long long* __restrict__ ckey_prev,
                             const float4* __restrict__ imp_prev, unsigned int n_prev, unsigned long long* keys,
                             float4* imps, size_t cap) {
    if (*need_hash == 0u) return;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_prev; i += gridDim.x * blockDim.x) {
        const unsigned long long key = ckey_prev[i];
        if (key != 0ull) ht_insert(keys, imps, cap, key, imp_prev[i]);
    }
}

// ---------------------------------------------------------------- contacts
// What the three rows of a contact need from its manifold.
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
__global__ void __launch_bounds__(TPB) k_assemble_contacts(
    unsigned int nC, unsigned int nJ, unsigned int maxc, const nb2_manifold* __restrict__ manifolds,
    const nb2_contact* __restrict__ contacts, const unsigned int* __restrict__ c_manifold,
    const unsigned int* __restrict__ chunk_base, BodyArrays B, SchedView vs, SchedView ps, RowOut out, float4* p_row,
    size_t n_pslots_max, ImpulseCacheView cache, float warmstart_coeff, float restitution_threshold, float inv_dt,
    int model) {
    const unsigned int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nC) return;
    const unsigned int m = c_manifold[ci];
    if (m == 0xFFFFFFFFu) return;
    const nb2_manifold& mf = manifolds[m];
    ContactQuads cq;
    load_contact(contacts, ci, &cq);
    const nb2_contact& c = *reinterpret_cast<const nb2_contact*>(&cq);
    BodySide s1, s2;
    load_side(B, mf.body1, &s1);
    load_side(B, mf.body2, &s2);
    if (s1.status != NB2_BODY_DYNAMIC && s2.status != NB2_BODY_DYNAMIC) return;  // filtered by the caller in the reference (mechanical_world.rs:287-300)
    const unsigned int lc = ci - mf.first_contact;
    const unsigned int chunk = chunk_base[m] + lc / NB2_CHUNK;
    const int lcc = (int)(lc % NB2_CHUNK);
    float4 cached = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c.key != 0ull) {  // impulse cache (signorini_coulomb_pyramid_model.rs:104-108)
        float4 prev;
        if (cache_fast_path(cache, ci, c.key, &prev) || cache_chunk_path(cache, mf.first_contact + NB2_CHUNK * (lc / NB2_CHUNK), ci, c.key, &prev))
            cached = prev;
        else if (!cache.chunk_local_ids) *cache.need_hash = 1u;
    }
    const size_t item_f = (size_t)nJ + chunk, item_n = (size_t)nJ + maxc + chunk;
    ContactSlots S;
    S.t1 = S.t2 = 0;
    if (model != NB2_CONTACT_SIGNORINI) {  // the frictionless model schedules no friction group
        S.t1 = vs.row_slot(item_f, 2 * lcc);
        S.t2 = vs.row_slot(item_f, 2 * lcc + 1);
    }
    S.n = vs.row_slot(item_n, lcc);
    S.p = ps.pos_slot((size_t)nJ + chunk, lcc);
    assemble_contact(out, s1, s2, manifold_consts(mf), c, f4_quat(B.pos_q[mf.body1]), cached, S, false, nullptr, p_row,
                     n_pslots_max, warmstart_coeff, restitution_threshold, inv_dt, model);
}

// ---------------------------------------------------------------- joints
// Warm-start slot of row r inside nb2_joint.impulses (what each velocity_constraints passes as
// `impulses[i]`), and the slot cache_impulses stores impulse_id == r into (verbatim, including
// the pin-slot / cylindrical quirk where id 2 lands in lin_impulses[2]).
__device__ __forceinline__ int joint_warm_slot(unsigned int type, int r) {
    switch (type) {
        case NB2_JOINT_BALL: return r;
        case NB2_JOINT_REVOLUTE:
        case NB2_JOINT_FIXED: return r;  // lin[r] | ang[r-3] == impulses[r]
        case NB2_JOINT_PRISMATIC: return r < 2 ? r : (r < 5 ? 3 + (r - 2) : 6);
        case NB2_JOINT_UNIVERSAL: return r;
        case NB2_JOINT_PLANAR:
        case NB2_JOINT_RECTANGULAR: return r == 0 ? 0 : 3 + (r - 1);
        case NB2_JOINT_PIN_SLOT:
        case NB2_JOINT_CYLINDRICAL: return r < 2 ? r : 3 + (r - 2);
        case NB2_JOINT_CARTESIAN: return 3 + r;
        default: return 0;
    }
}
__device__ __forceinline__ int joint_cache_slot(unsigned int type, int id) {
    switch (type) {
        case NB2_JOINT_BALL: return id;
        case NB2_JOINT_REVOLUTE:
        case NB2_JOINT_PIN_SLOT:
        case NB2_JOINT_CYLINDRICAL:
        case NB2_JOINT_FIXED: return id;  // id<3 -> lin[id]; else ang[id-3]
        case NB2_JOINT_PRISMATIC: return id < 2 ? id : (id < 5 ? 3 + (id + 1 - 3) : 6);
        case NB2_JOINT_UNIVERSAL: return id < 3 ? id : 3;
        case NB2_JOINT_PLANAR:
        case NB2_JOINT_RECTANGULAR: return id == 0 ? 0 : 3 + (id - 1);
        case NB2_JOINT_CARTESIAN: return 3 + id;
        default: return 0;
    }
}

struct JointFrame {
    Pose pos1, pos2;  // position_at_material_point (* ref_frame for fixed/cartesian)
};
__device__ __forceinline__ JointFrame joint_frames(const nb2_joint& j, const Pose& p1, const Pose& p2) {
    JointFrame f;
    f.pos1.t = p1.t + quat_rotate(p1.r, mk3(j.anchor1[0], j.anchor1[1], j.anchor1[2]));
    f.pos1.r = p1.r;
    f.pos2.t = p2.t + quat_rotate(p2.r, mk3(j.anchor2[0], j.anchor2[1], j.anchor2[2]));
    f.pos2.r = p2.r;
    if (j.type == NB2_JOINT_FIXED || j.type == NB2_JOINT_CARTESIAN) {
        f.pos1.r = quat_mul(f.pos1.r, mkq(j.ref_frame1[0], j.ref_frame1[1], j.ref_frame1[2], j.ref_frame1[3]));
        f.pos2.r = quat_mul(f.pos2.r, mkq(j.ref_frame2[0], j.ref_frame2[1], j.ref_frame2[2], j.ref_frame2[3]));
    }
    return f;
}

__global__ void __launch_bounds__(TPB) k_assemble_joints(unsigned int nJ, const nb2_joint* __restrict__ joints,
                                                         BodyArrays B, SchedView vs, RowOut out) {
    unsigned int ji = blockIdx.x * blockDim.x + threadIdx.x;
    if (ji >= nJ) return;
    if (vs.it_type[ji] == NB2_ITEM_INVALID) return;
    const nb2_joint& j = joints[ji];
    BodySide s1, s2;
    load_side(B, j.body1, &s1);
    load_side(B, j.body2, &s2);
    Pose p1, p2;
    p1.t = f4_xyz(B.pos_t[j.body1]);
    p1.r = f4_quat(B.pos_q[j.body1]);
    p2.t = f4_xyz(B.pos_t[j.body2]);
    p2.r = f4_quat(B.pos_q[j.body2]);
    JointFrame fr = joint_frames(j, p1, p2);
    const Vec3 a1 = fr.pos1.t, a2 = fr.pos2.t;
    const Vec3 ex = mk3(1.f, 0.f, 0.f), ey = mk3(0.f, 1.f, 0.f), ez = mk3(0.f, 0.f, 1.f);

    int r = 0;
    float J1[6], J2[6], W1[6], W2[6], rhs, rr;
    auto row = [&](bool angular, Vec3 dir, float lo) {
        emit_pair_row(out, 0, s1, s2, a1, a2, angular, dir, 0.f, &rhs, &rr, J1, J2, W1, W2);
        float warm = j.impulses[joint_warm_slot(j.type, r)];
        write_row(out, vs.row_slot(ji, r), J1, J2, W1, W2, rhs, rr, lo, NB2_F32_MAX, NB2_ROW_BILATERAL, 0, warm);
        ++r;
    };
    auto skip = [&]() {
        // a reserved but unused row: every plane is written.  The staged kernel streams and multiplies all
        // of them unconditionally, and 0 * the stale NaN/Inf bits of a recycled allocation would poison
        // mj_lambda (tests/test_gpu_parity.py::test_unused_joint_rows_ignore_stale_memory)
        const size_t slot = vs.row_slot(ji, r);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) out.jac[(size_t)k * out.n_slots_max + slot] = z;
        out.jac[4 * out.n_slots_max + slot] = make_float4(0.f, 0.f, __int_as_float(NB2_ROW_NONE), 0.f);
        out.hdr[slot] = z;
        out.imp[slot] = 0.f;
        ++r;
    };
    auto lin3 = [&]() {  // helper::cancel_relative_linear_velocity (helper.rs:247-293)
        row(false, ex, -NB2_F32_MAX);
        row(false, ey, -NB2_F32_MAX);
        row(false, ez, -NB2_F32_MAX);
    };
    auto ang3 = [&]() {  // helper::cancel_relative_angular_velocity (helper.rs:502-548)
        row(true, ex, -NB2_F32_MAX);
        row(true, ey, -NB2_F32_MAX);
        row(true, ez, -NB2_F32_MAX);
    };
    auto restrict2 = [&](bool angular, Vec3 axis) {  // helper.rs:614-695 / 771-853
        Vec3 t1, t2;
        tangent_basis(axis, &t1, &t2);
        row(angular, t1, -NB2_F32_MAX);
        row(angular, t2, -NB2_F32_MAX);
    };
    const Vec3 ax1 = mk3(j.axis1[0], j.axis1[1], j.axis1[2]);
    const Vec3 ax2 = mk3(j.axis2[0], j.axis2[1], j.axis2[2]);
    const Vec3 ax3 = mk3(j.axis3[0], j.axis3[1], j.axis3[2]);
    switch (j.type) {
        case NB2_JOINT_BALL:  // ball_constraint.rs:92-126
            lin3();
            break;
        case NB2_JOINT_REVOLUTE:  // revolute_constraint.rs:194-258
            lin3();
            restrict2(true, quat_rotate(fr.pos1.r, ax1));
            break;
        case NB2_JOINT_PRISMATIC: {  // prismatic_constraint.rs:149-236 + unit_constraint.rs:10-125
            Vec3 axis = quat_rotate(fr.pos1.r, ax1);
            restrict2(false, axis);
            ang3();
            const bool has_min = j.flags & NB2_JOINT_FLAG_MIN_OFFSET, has_max = j.flags & NB2_JOINT_FLAG_MAX_OFFSET;
            float offset = dot3(axis, a2 - a1);
            int act = 0;  // 0 none, 1 bilateral +axis, 2 unilateral -axis, 3 unilateral +axis
            if (has_min && has_max) {
                float diff = fabsf(j.min_offset - j.max_offset);
                float largest = fmaxf(fabsf(j.min_offset), fabsf(j.max_offset));
                bool eq = (j.min_offset == j.max_offset) || diff <= NB2_F32_EPS || diff <= largest * NB2_F32_EPS;
                if (eq) act = 1;
                else if (offset <= j.min_offset) act = 2;
                else if (offset >= j.max_offset) act = 3;
            } else if (has_min) {
                if (offset <= j.min_offset) act = 2;
            } else if (has_max) {
                if (offset >= j.max_offset) act = 3;
            }
            if (act == 0) skip();
            else row(false, act == 2 ? -axis : axis, act == 1 ? -NB2_F32_MAX : 0.f);
            skip();  // the reference reserves 7 rows (prismatic_constraint.rs:124-126)
            break;
        }
        case NB2_JOINT_UNIVERSAL: {  // universal_constraint.rs:104-165
            lin3();
            Vec3 axis1 = quat_rotate(fr.pos1.r, ax1), axis2 = quat_rotate(fr.pos2.r, ax2);
            Vec3 orth;
            float len;
            if (unit_try_new_and_get(cross3(axis1, axis2), NB2_F32_EPS, &orth, &len)) row(true, orth, -NB2_F32_MAX);
            else skip();
            break;
        }
        case NB2_JOINT_PLANAR: {  // planar_constraint.rs:101-163
            Vec3 axis1 = quat_rotate(fr.pos1.r, ax1);
            row(false, axis1, -NB2_F32_MAX);
            restrict2(true, axis1);
            break;
        }
        case NB2_JOINT_RECTANGULAR: {  // rectangular_constraint.rs:99-160
            row(false, quat_rotate(fr.pos1.r, ax1), -NB2_F32_MAX);
            ang3();
            break;
        }
        case NB2_JOINT_PIN_SLOT:  // pin_slot_constraint.rs:149-212
            restrict2(false, quat_rotate(fr.pos1.r, ax1));
            restrict2(true, quat_rotate(fr.pos1.r, ax3));
            break;
        case NB2_JOINT_CYLINDRICAL: {  // cylindrical_constraint.rs:143-205
            Vec3 axis1 = quat_rotate(fr.pos1.r, ax1);
            restrict2(false, axis1);
            restrict2(true, axis1);
            break;
        }
        case NB2_JOINT_FIXED:  // fixed_constraint.rs:113-171
            lin3();
            ang3();
            break;
        case NB2_JOINT_CARTESIAN:  // cartesian_constraint.rs:106-144
            ang3();
            break;
        default:
            break;
    }
    (void)ax2;
}

// ---------------------------------------------------------------- cache impulses
__global__ void __launch_bounds__(TPB) k_cache_contact_impulses(
    int mode, unsigned int nC, unsigned int nJ, unsigned int maxc, const nb2_manifold* __restrict__ manifolds,
    const nb2_contact* __restrict__ contacts, const unsigned int* __restrict__ c_manifold,
    const unsigned int* __restrict__ chunk_base, const int* __restrict__ status, SchedView vs,
    const float* __restrict__ r_imp, const float4* __restrict__ c_geo, size_t n_pslots_max, float4* imp_cur,
    unsigned long long* ckey_cur, int compact_layout, int model, ImpulseCacheView prev) {
    unsigned int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nC) return;
    const unsigned int m = c_manifold[ci];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned long long key = contacts[ci].key;
    ckey_cur[ci] = key;
    if (m == 0xFFFFFFFFu) {
        imp_cur[ci] = v;
        return;
    }
    const nb2_manifold& mf = manifolds[m];
    if (status[mf.body1] != NB2_BODY_DYNAMIC && status[mf.body2] != NB2_BODY_DYNAMIC) {
        // Reported but not solved this step: the pair sleeps (it is filtered out of the manifold list,
        // mechanical_world.rs:287-300).  The reference's cache never forgets, so the entry is carried over and
        // the island is warm-started when it wakes: from the per-contact arrays when the contact kept its
        // index, else from the hash table if this step's assembly happened to build it.
        if (key != 0ull && !cache_fast_path(prev, ci, key, &v) &&
            !cache_chunk_path(prev, mf.first_contact + NB2_CHUNK * ((ci - mf.first_contact) / NB2_CHUNK), ci, key, &v)) {
            if (!(*prev.need_hash != 0u && ht_lookup(prev.ht_keys, prev.ht_imps, prev.ht_cap, key, &v)))
                v = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        imp_cur[ci] = v;
        return;
    }
    {
        const unsigned int lc = ci - mf.first_contact;
        const unsigned int chunk = chunk_base[m] + lc / NB2_CHUNK;
        const int lcc = (int)(lc % NB2_CHUNK);
        const int ncc = min(NB2_CHUNK, (int)mf.num_contacts - NB2_CHUNK * (int)(lc / NB2_CHUNK));
        if (compact_layout) {
            const float4 q = c_geo[4 * n_pslots_max + vs.pos_slot((size_t)nJ + chunk, lcc)];
            v = make_float4(q.x, q.y, q.z, 0.f);
        } else if (model == NB2_CONTACT_SIGNORINI) {  // one row per contact (an inactive contact's slot carries its old impulse)
            v.x = r_imp[mode == NB2_MODE_COLOURED ? vs.row_slot((size_t)nJ + chunk, lcc)
                                                  : vs.row_slot((size_t)nJ + maxc + chunk, lcc)];
        } else if (mode == NB2_MODE_COLOURED) {
            size_t item = (size_t)nJ + chunk;
            v.x = r_imp[vs.row_slot(item, 2 * ncc + lcc)];
            v.y = r_imp[vs.row_slot(item, 2 * lcc)];
            v.z = r_imp[vs.row_slot(item, 2 * lcc + 1)];
        } else {
            size_t item_f = (size_t)nJ + chunk, item_n = (size_t)nJ + maxc + chunk;
            v.x = r_imp[vs.row_slot(item_n, lcc)];
            v.y = r_imp[vs.row_slot(item_f, 2 * lcc)];
            v.z = r_imp[vs.row_slot(item_f, 2 * lcc + 1)];
        }
    }
    imp_cur[ci] = v;
}

// Steps on which some contact missed the fast path of the impulse cache: the table over the previous step's
// keys has been built by now; look the missed contacts up and patch their warm-start impulses in place
// (assembly wrote zeros for them).  Same slot arithmetic as k_cache_contact_impulses.
__global__ void __launch_bounds__(TPB) k_warm_fixup(
    int mode, unsigned int nC, unsigned int nJ, unsigned int maxc, const nb2_manifold* __restrict__ manifolds,
    const nb2_contact* __restrict__ contacts, const unsigned int* __restrict__ c_manifold,
    const unsigned int* __restrict__ chunk_base, const int* __restrict__ status, SchedView vs, float* r_imp,
    float4* c_geo, size_t n_pslots_max, ImpulseCacheView cache, float warmstart_coeff, int compact_layout, int model,
    const float4* __restrict__ r_kind /* jac plane 4 */) {
    if (*cache.need_hash == 0u) return;
    unsigned int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nC) return;
    const unsigned int m = c_manifold[ci];
    if (m == 0xFFFFFFFFu) return;
    const unsigned long long key = contacts[ci].key;
    if (key == 0ull) return;
    float4 v;
    const nb2_manifold& mf = manifolds[m];
    if (cache_fast_path(cache, ci, key, &v)) return;  // already warm-started by the assembly
    if (cache_chunk_path(cache, mf.first_contact + NB2_CHUNK * ((ci - mf.first_contact) / NB2_CHUNK), ci, key, &v)) return;
    if (!ht_lookup(cache.ht_keys, cache.ht_imps, cache.ht_cap, key, &v)) return;
    if (status[mf.body1] != NB2_BODY_DYNAMIC && status[mf.body2] != NB2_BODY_DYNAMIC) return;
    const unsigned int lc = ci - mf.first_contact;
    const unsigned int chunk = chunk_base[m] + lc / NB2_CHUNK;
    const int lcc = (int)(lc % NB2_CHUNK);
    const int ncc = min(NB2_CHUNK, (int)mf.num_contacts - NB2_CHUNK * (int)(lc / NB2_CHUNK));
    const float wn = v.x * warmstart_coeff, wt1 = v.y * warmstart_coeff, wt2 = v.z * warmstart_coeff;
    if (compact_layout) {
        c_geo[4 * n_pslots_max + vs.pos_slot((size_t)nJ + chunk, lcc)] = make_float4(wn, wt1, wt2, 1.f);
    } else if (model == NB2_CONTACT_SIGNORINI) {
        const size_t slot = mode == NB2_MODE_COLOURED ? vs.row_slot((size_t)nJ + chunk, lcc) : vs.row_slot((size_t)nJ + maxc + chunk, lcc);
        // an inactive contact (NONE row) carries the raw cached impulse, an active one is warm-started
        r_imp[slot] = __float_as_int(r_kind[slot].z) == NB2_ROW_NONE ? v.x : wn;
    } else if (mode == NB2_MODE_COLOURED) {
        size_t item = (size_t)nJ + chunk;
        r_imp[vs.row_slot(item, 2 * ncc + lcc)] = wn;
        r_imp[vs.row_slot(item, 2 * lcc)] = wt1;
        r_imp[vs.row_slot(item, 2 * lcc + 1)] = wt2;
    } else {
        size_t item_f = (size_t)nJ + chunk, item_n = (size_t)nJ + maxc + chunk;
        r_imp[vs.row_slot(item_n, lcc)] = wn;
        r_imp[vs.row_slot(item_f, 2 * lcc)] = wt1;
        r_imp[vs.row_slot(item_f, 2 * lcc + 1)] = wt2;
    }
}

__global__ void __launch_bounds__(TPB) k_cache_joint_impulses(unsigned int nJ, nb2_joint* joints, SchedView vs,
                                                              const int* __restrict__ it_nrows,
                                                              const float4* __restrict__ r_kind /* jac plane 4 */,
                                                              const float* __restrict__ r_imp, float inv_dt) {
    unsigned int ji = blockIdx.x * blockDim.x + threadIdx.x;
    if (ji >= nJ) return;
    if (vs.it_type[ji] == NB2_ITEM_INVALID) return;
    nb2_joint& j = joints[ji];
    const int nrows = it_nrows[ji];
    for (int r = 0; r < nrows; ++r) {
        size_t slot = vs.row_slot(ji, r);
        if (__float_as_int(r_kind[slot].z) == NB2_ROW_NONE) continue;
        j.impulses[joint_cache_slot(j.type, r)] = r_imp[slot];
    }
    const float* lin = &j.impulses[0];
    const float* ang = &j.impulses[3];
    const float inv_dt2 = inv_dt * inv_dt;
    const float lin_sq = norm_sq3(mk3(lin[0], lin[1], lin[2]));
    const float ang_sq = norm_sq3(mk3(ang[0], ang[1], ang[2]));
    bool broken;
    switch (j.type) {
        case NB2_JOINT_BALL: broken = lin_sq * inv_dt * inv_dt > j.break_force_squared; break;
        case NB2_JOINT_UNIVERSAL:
            broken = lin_sq * inv_dt2 > j.break_force_squared || ang[0] * ang[0] * inv_dt2 > j.break_torque_squared;
            break;
        case NB2_JOINT_PLANAR:
            broken = lin[0] * lin[0] * inv_dt2 > j.break_force_squared ||
                     ang[0] * ang[0] * inv_dt2 + ang[1] * ang[1] * inv_dt2 > j.break_torque_squared;
            break;
        case NB2_JOINT_RECTANGULAR:
            broken = lin[0] * lin[0] * inv_dt2 > j.break_force_squared || ang_sq * inv_dt2 > j.break_torque_squared;
            break;
        case NB2_JOINT_CARTESIAN: broken = ang_sq * inv_dt * inv_dt > j.break_torque_squared; break;
        default:
            broken = lin_sq * inv_dt2 > j.break_force_squared || ang_sq * inv_dt2 > j.break_torque_squared;
            break;
    }
    if (broken) j.broken = 1u;
}

static BodyArrays body_arrays(Context* ctx) {
    BodyArrays B;
    B.raw = ctx->raw.p;
    B.pos_t = ctx->pos_t.p;
    B.pos_q = ctx->pos_q.p;
    B.vel = ctx->vel.p;
    B.com_im = ctx->com_im.p;
    B.inv_i = ctx->inv_i.p;
    B.ext = ctx->ext.p;
    return B;
}
static RowOut row_out(Context* ctx) {
    RowOut o;
    o.jac = ctx->r_jac.p;
    o.hdr = ctx->r_hdr.p;
    o.imp = ctx->r_imp.p;
    o.n_slots_max = ctx->n_slots_max;
    return o;
}

int launch_assemble(Context* ctx, int mode) {
    const bool ref = mode == NB2_MODE_REFERENCE_ORDER;
    const size_t maxc = ctx->max_chunks;
    // row-slot upper bounds (ELL padding included): every item may be padded to the widest group
    const size_t n_items = ctx->vs.n_items;
    // (coloured mode: only joints own generic rows, contacts live in the compact c_geo planes)
    const size_t slots = (ctx->step_layout == 0 ? n_items * (size_t)(3 * NB2_CHUNK)
                                                : (ctx->n_joints ? n_items * (size_t)NB2_MAX_JOINT_ROWS : 0)) + 16;
    const size_t pitems = ref ? ctx->ps.n_items : n_items;
    const size_t pslots = pitems * (size_t)NB2_CHUNK + 16;
    ctx->n_slots_max = slots;
    ctx->n_pslots_max = pslots;
    NB2_TRY(ctx->r_jac.reserve(ctx, NB2_ROW_PLANES * slots));
    ctx->n_slots_max = ctx->r_jac.cap / NB2_ROW_PLANES;  // keep the plane stride consistent with the allocation
    NB2_TRY(ctx->r_hdr.reserve(ctx, ctx->n_slots_max));
    NB2_TRY(ctx->r_imp.reserve(ctx, ctx->n_slots_max));
    NB2_TRY(ctx->p_row.reserve(ctx, 5 * pslots));
    ctx->n_pslots_max = ctx->p_row.cap / 5;
    NB2_TRY(ctx->c_geo.reserve(ctx, ctx->step_layout == 0 ? 16 : 5 * ctx->n_pslots_max));
    NB2_TRY(ctx->p_hdr.reserve(ctx, ref ? 16 : 5 * (n_items + 16)));
    ctx->n_ghdr_max = ctx->p_hdr.cap / 5;
    SchedView vs = view_of(ctx->vs);
    SchedView ps = ref ? view_of(ctx->ps) : vs;
    const int prev = 1 - ctx->cur;
    if (ctx->poison_rows) {  // every byte the solve kernels read must have been written by this step's assembly
        NB2_CUDA(ctx, cudaMemsetAsync(ctx->r_jac.p, 0xFF, ctx->r_jac.cap * sizeof(float4), ctx->stream));
        NB2_CUDA(ctx, cudaMemsetAsync(ctx->r_hdr.p, 0xFF, ctx->r_hdr.cap * sizeof(float4), ctx->stream));
        NB2_CUDA(ctx, cudaMemsetAsync(ctx->r_imp.p, 0xFF, ctx->r_imp.cap * sizeof(float), ctx->stream));
    }
    if (ctx->n_joints) {
        k_assemble_joints<<<nblk(ctx->n_joints), TPB, 0, ctx->stream>>>(ctx->n_joints, ctx->joints.p,
                                                                        body_arrays(ctx), vs, row_out(ctx));
        ctx->launches++;
    }
    if (ctx->n_contacts) {
        // impulse cache of the previous step: per-contact arrays first, hash table only if some contact misses
        NB2_TRY(ctx->flags.reserve(ctx, 4));
        unsigned int* need_hash = ctx->flags.p + 2;
        ImpulseCacheView cache;
        cache.ckey_prev = ctx->ckey[prev].p;
        cache.imp_prev = ctx->imp[prev].p;
        cache.n_prev = ctx->imp_n[prev];
        cache.ht_keys = ctx->ht_keys[prev].p;
        cache.ht_imps = ctx->ht_imps[prev].p;
        cache.ht_cap = ctx->ht_cap[prev];
        cache.need_hash = need_hash;
        cache.chunk_local_ids = ctx->manifolds_from_producer ? 1 : 0;
        NB2_CUDA(ctx, cudaMemsetAsync(need_hash, 0, sizeof(unsigned int), ctx->stream));
        if (ref) {
            k_assemble_contacts<<<nblk(ctx->n_contacts), TPB, 0, ctx->stream>>>(
                ctx->n_contacts, ctx->n_joints, (unsigned int)maxc, ctx->manifolds.p, ctx->contacts.p, ctx->c_manifold.p,
                ctx->chunk_base.p, body_arrays(ctx), vs, ps, row_out(ctx), ctx->p_row.p, ctx->n_pslots_max, cache,
                ctx->params.warmstart_coeff, ctx->params.restitution_velocity_threshold, ctx->inv_dt, ctx->contact_model);
        } else {  // one thread per group slot in ELL order
            k_assemble_groups<<<nblk(n_items), TPB, 0, ctx->stream>>>(
                ctx->n_contacts, ctx->manifolds.p, ctx->contacts.p, ctx->chunk_base.p, ctx->chunk_manifold.p,
                body_arrays(ctx), vs, row_out(ctx), ctx->p_row.p, ctx->n_pslots_max, ctx->p_hdr.p, ctx->n_ghdr_max,
                ctx->c_geo.p, cache, ctx->params.warmstart_coeff, ctx->params.restitution_velocity_threshold, ctx->inv_dt,
                ctx->step_layout, ctx->contact_model);
        }
        ctx->launches++;
        if (cache.ht_cap) {  // all three return at once unless a contact missed the fast path
            const unsigned int few = (unsigned int)ctx->sm_count * 16;  // grid-stride: the common case is an immediate return
            k_hash_clear<<<min(nblk(cache.ht_cap), few), TPB, 0, ctx->stream>>>(need_hash, ctx->ht_keys[prev].p, cache.ht_cap);
            k_hash_build<<<min(nblk(cache.n_prev), few), TPB, 0, ctx->stream>>>(need_hash, cache.ckey_prev, cache.imp_prev, cache.n_prev,
                                                                      ctx->ht_keys[prev].p, ctx->ht_imps[prev].p, cache.ht_cap);
            k_warm_fixup<<<nblk(ctx->n_contacts), TPB, 0, ctx->stream>>>(
                mode, ctx->n_contacts, ctx->n_joints, (unsigned int)maxc, ctx->manifolds.p, ctx->contacts.p, ctx->c_manifold.p,
                ctx->chunk_base.p, ctx->b_status.p, vs, ctx->r_imp.p, ctx->c_geo.p, ctx->n_pslots_max, cache,
                ctx->params.warmstart_coeff, ctx->step_layout, ctx->contact_model, ctx->r_jac.p + 4 * ctx->n_slots_max);
            ctx->launches += 3;
        }
    }
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int launch_cache_impulses(Context* ctx, int mode) {
    const int cur = ctx->cur;
    // size the table for this step's contacts
    size_t cap = 0;
    if (ctx->n_contacts) {
        cap = 64;
        while (cap < 2 * (size_t)ctx->n_contacts) cap <<= 1;
    }
    NB2_TRY(ctx->imp[cur].reserve(ctx, ctx->n_contacts + 1));
    NB2_TRY(ctx->ckey[cur].reserve(ctx, ctx->n_contacts + 1));
    if (cap) {  // the table itself is filled lazily by the next step's assembly (k_hash_clear / k_hash_build)
        NB2_TRY(ctx->ht_keys[cur].reserve(ctx, cap));
        NB2_TRY(ctx->ht_imps[cur].reserve(ctx, cap));
    }
    ctx->ht_cap[cur] = cap;
    ctx->imp_n[cur] = ctx->n_contacts;
    SchedView vs = view_of(ctx->vs);
    if (ctx->n_contacts) {
        const int prev = 1 - cur;
        ImpulseCacheView pv;
        pv.ckey_prev = ctx->ckey[prev].p;
        pv.imp_prev = ctx->imp[prev].p;
        pv.n_prev = ctx->imp_n[prev];
        pv.ht_keys = ctx->ht_keys[prev].p;
        pv.ht_imps = ctx->ht_imps[prev].p;
        pv.ht_cap = ctx->ht_cap[prev];
        pv.need_hash = ctx->flags.p + 2;
        pv.chunk_local_ids = ctx->manifolds_from_producer ? 1 : 0;
        k_cache_contact_impulses<<<nblk(ctx->n_contacts), TPB, 0, ctx->stream>>>(
            mode, ctx->n_contacts, ctx->n_joints, (unsigned int)ctx->max_chunks, ctx->manifolds.p, ctx->contacts.p,
            ctx->c_manifold.p, ctx->chunk_base.p, ctx->b_status.p, vs, ctx->r_imp.p, ctx->c_geo.p, ctx->n_pslots_max,
            ctx->imp[cur].p, ctx->ckey[cur].p, ctx->step_layout, ctx->contact_model, pv);
        ctx->launches++;
    }
    if (ctx->n_joints) {
        k_cache_joint_impulses<<<nblk(ctx->n_joints), TPB, 0, ctx->stream>>>(
            ctx->n_joints, ctx->joints.p, vs, ctx->vs.it_nrows.p, ctx->r_jac.p + 4 * ctx->n_slots_max, ctx->r_imp.p, ctx->inv_dt);
        ctx->launches++;
    }
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

}  // namespace nb2
