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
#include "assemble_contact.cuh"

namespace nb2 {

static const int TPB = ASM_TPB;
static inline unsigned int nblk(size_t n) { return (unsigned int)((n + TPB - 1) / TPB); }

int launch_assemble_groups(Context* ctx, size_t n_items, const BodyArrays& B, const SchedView& vs, const RowOut& out,
                           const ImpulseCacheView& cache);  // assemble_coloured.cu

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
            NB2_TRY(launch_assemble_groups(ctx, n_items, body_arrays(ctx), vs, row_out(ctx), cache));
            ctx->launches--;  // counted below
        }
        ctx->launches++;
        // the producer's contacts are looked up inside their own pair (cache_chunk_path) and never ask for the table
        if (cache.ht_cap && !cache.chunk_local_ids) {  // all three return at once unless a contact missed the fast path
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
