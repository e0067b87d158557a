// Device helpers of the position solve shared by the reference-order kernel (solve.cu) and the
// coloured staged kernel (solve_coloured.cu): body state, GenericNonlinearConstraint resolution
// (nonlinear_sor_prox.rs:78-119), the joints' position constraints and ncollide's
// ContactKinematic::contact.
#pragma once
#include "solve_common.cuh"

namespace nb2 {

struct PosBody {
    BodyPose bp;
    float inv_mass;
    Mat3 inv_i;
    float mask[6];
    Vec3 local_com;
    bool dynamic;
};
struct PosArrays {
    const nb2_body* raw;
    PoseQuads pos_t;
    PoseQuads pos_q;
    float4* com_im;
    const float4* inv_i;
};
__device__ __forceinline__ void load_pos_body(const PosArrays& A, int idx, PosBody* o) {
    // three quads of the 176-byte record instead of ten scalar loads: local_com (bytes 52..63),
    // jacobian_mask[6] + status (bytes 144..171)
    const float4* rq = reinterpret_cast<const float4*>(&A.raw[idx]);
    const float4 q3 = __ldg(rq + 3), m0 = __ldg(rq + 9), m1 = __ldg(rq + 10);
    o->dynamic = __float_as_int(m1.z) == NB2_BODY_DYNAMIC;
    o->bp.pose.t = f4_xyz(ldcg4(&A.pos_t[idx]));
    o->bp.pose.r = f4_quat(ldcg4(&A.pos_q[idx]));
    float4 c = ldcg4(&A.com_im[idx]);
    o->bp.com = f4_xyz(c);
    o->inv_mass = c.w;
    float4 r0 = A.inv_i[3 * idx], r1 = A.inv_i[3 * idx + 1], r2 = A.inv_i[3 * idx + 2];
    o->inv_i.m[0][0] = r0.x; o->inv_i.m[0][1] = r0.y; o->inv_i.m[0][2] = r0.z;
    o->inv_i.m[1][0] = r1.x; o->inv_i.m[1][1] = r1.y; o->inv_i.m[1][2] = r1.z;
    o->inv_i.m[2][0] = r2.x; o->inv_i.m[2][1] = r2.y; o->inv_i.m[2][2] = r2.z;
    o->mask[0] = m0.x; o->mask[1] = m0.y; o->mask[2] = m0.z; o->mask[3] = m0.w; o->mask[4] = m1.x; o->mask[5] = m1.y;
    o->local_com = mk3(q3.y, q3.z, q3.w);
}
__device__ __forceinline__ void store_pos_body(const PosArrays& A, int idx, const PosBody& b) {
    stcg4(&A.pos_t[idx], xyz_f4(b.bp.pose.t, 0.f));
    stcg4(&A.pos_q[idx], quat_f4(b.bp.pose.r));
    stcg4(&A.com_im[idx], xyz_f4(b.bp.com, b.inv_mass));
}
// fill_constraint_geometry without velocities: weighted jacobian + inv_r contribution
__device__ __forceinline__ void pos_fill(const PosBody& b, Vec3 point, bool angular, Vec3 dir, Vec3* wl, Vec3* wa,
                                         float* inv_r) {
    if (!b.dynamic) return;
    Vec3 pos = point - b.bp.com;
    Vec3 fl = angular ? mk3(0.f, 0.f, 0.f) : dir;
    Vec3 fa = angular ? dir : cross3(pos, dir);
    Vec3 ml = mul3(fl, mk3(b.mask[0], b.mask[1], b.mask[2]));
    Vec3 ma = mul3(fa, mk3(b.mask[3], b.mask[4], b.mask[5]));
    *wl = ml * b.inv_mass;
    *wa = mat_vec(b.inv_i, ma);
    *inv_r += b.inv_mass + dot3(ma, *wa);
}
struct PosParams {
    float erp, allowed_lin, allowed_ang, max_lin, max_ang;
};
__device__ __forceinline__ float clamp_rhs(float rhs, bool angular, const PosParams& P) {
    if (angular) return fmaxf((rhs + P.allowed_ang) * P.erp, -P.max_ang);
    return fmaxf((rhs + P.allowed_lin) * P.erp, -P.max_lin);
}
// GenericNonlinearConstraint resolution (helper position generators + solve_generic,
// nonlinear_sor_prox.rs:78-119)
__device__ __forceinline__ void solve_generic(PosBody* b1, PosBody* b2, Vec3 a1, Vec3 a2, bool angular, Vec3 dir,
                                              float rhs_in, const PosParams& P) {
    Vec3 w1l = mk3(0.f, 0.f, 0.f), w1a = w1l, w2l = w1l, w2a = w1l;
    float inv_r = 0.f;
    pos_fill(*b1, a1, angular, dir, &w1l, &w1a, &inv_r);
    pos_fill(*b2, a2, angular, -dir, &w2l, &w2a, &inv_r);
    float r = inv_r != 0.f ? 1.f / inv_r : 1.f;
    float rhs = clamp_rhs(rhs_in, angular, P);
    if (rhs < 0.f) {
        float impulse = -rhs * r;
        if (b1->dynamic) apply_displacement(&b1->bp, b1->local_com, w1l * impulse, w1a * impulse);
        if (b2->dynamic) apply_displacement(&b2->bp, b2->local_com, w2l * impulse, w2a * impulse);
    }
}
__device__ __forceinline__ Vec3 pi_fallback_axis(Vec3 axis1) {  // helper.rs:720-724
    int imin = 0;
    float best = fabsf(axis1.x);
    if (fabsf(axis1.y) < best) {
        best = fabsf(axis1.y);
        imin = 1;
    }
    if (fabsf(axis1.z) < best) imin = 2;
    Vec3 e = mk3(imin == 0 ? 1.f : 0.f, imin == 1 ? 1.f : 0.f, imin == 2 ? 1.f : 0.f);
    Vec3 c = cross3(e, axis1);
    return (c / norm3(c)) * NB2_PI;
}

// the position constraints of one joint, in the reference's order
static __device__ void joint_position(const nb2_joint& j, PosBody* b1, PosBody* b2, const PosParams& P) {
    const Vec3 ax1 = mk3(j.axis1[0], j.axis1[1], j.axis1[2]);
    const Vec3 ax2 = mk3(j.axis2[0], j.axis2[1], j.axis2[2]);
    const Vec3 ax3 = mk3(j.axis3[0], j.axis3[1], j.axis3[2]);
    int n;
    switch (j.type) {
        case NB2_JOINT_BALL:
        case NB2_JOINT_CARTESIAN: n = 1; break;
        case NB2_JOINT_PRISMATIC: n = (j.flags & (NB2_JOINT_FLAG_MIN_OFFSET | NB2_JOINT_FLAG_MAX_OFFSET)) ? 3 : 2; break;
        default: n = 2; break;
    }
    for (int i = 0; i < n; ++i) {
        Pose pos1, pos2;
        pos1.t = b1->bp.pose.t + quat_rotate(b1->bp.pose.r, mk3(j.anchor1[0], j.anchor1[1], j.anchor1[2]));
        pos1.r = b1->bp.pose.r;
        pos2.t = b2->bp.pose.t + quat_rotate(b2->bp.pose.r, mk3(j.anchor2[0], j.anchor2[1], j.anchor2[2]));
        pos2.r = b2->bp.pose.r;
        if (j.type == NB2_JOINT_FIXED || j.type == NB2_JOINT_CARTESIAN) {
            pos1.r = quat_mul(pos1.r, mkq(j.ref_frame1[0], j.ref_frame1[1], j.ref_frame1[2], j.ref_frame1[3]));
            pos2.r = quat_mul(pos2.r, mkq(j.ref_frame2[0], j.ref_frame2[1], j.ref_frame2[2], j.ref_frame2[3]));
        }
        const Vec3 a1 = pos1.t, a2 = pos2.t;
        // 0 translation, 1 align_axis, 2 rotation, 3 project, 4 limits, 5 restore angle, 6 translation wrt axis
        int what = -1;
        Vec3 u = mk3(0.f, 0.f, 0.f), v = u;
        switch (j.type) {
            case NB2_JOINT_BALL: what = 0; break;
            case NB2_JOINT_REVOLUTE:
                what = i == 0 ? 0 : 1;
                u = quat_rotate(pos1.r, ax1);
                v = quat_rotate(pos2.r, ax2);
                break;
            case NB2_JOINT_PRISMATIC:
                what = i == 0 ? 2 : (i == 1 ? 3 : 4);
                u = quat_rotate(pos1.r, ax1);
                break;
            case NB2_JOINT_UNIVERSAL:
                what = i == 0 ? 0 : 5;
                u = quat_rotate(pos1.r, ax1);
                v = quat_rotate(pos2.r, ax2);
                break;
            case NB2_JOINT_PLANAR:
                what = i == 0 ? 6 : 1;
                u = quat_rotate(pos1.r, ax1);
                v = quat_rotate(pos2.r, ax2);
                break;
            case NB2_JOINT_RECTANGULAR:
                what = i == 0 ? 6 : 2;
                u = quat_rotate(pos1.r, ax1);
                break;
            case NB2_JOINT_PIN_SLOT:
                what = i == 0 ? 1 : 3;
                if (i == 0) {
                    u = quat_rotate(pos1.r, ax3);
                    v = quat_rotate(pos2.r, ax2);
                } else {
                    u = quat_rotate(pos1.r, ax1);
                }
                break;
            case NB2_JOINT_CYLINDRICAL:
                what = i == 0 ? 1 : 3;
                u = quat_rotate(pos1.r, ax1);
                v = quat_rotate(pos2.r, ax2);
                break;
            case NB2_JOINT_FIXED: what = i == 0 ? 2 : 0; break;
            case NB2_JOINT_CARTESIAN: what = 2; break;
            default: break;
        }
        Vec3 dir;
        float depth;
        switch (what) {
            case 0:  // cancel_relative_translation (helper.rs:364-417)
                if (unit_try_new_and_get(a2 - a1, P.allowed_lin, &dir, &depth))
                    solve_generic(b1, b2, a1, a2, false, dir, -depth, P);
                break;
            case 1: {  // align_axis (helper.rs:701-766)
                Vec3 error;
                Quat rot;
                if (quat_rotation_between_axis(u, v, &rot)) error = quat_scaled_axis(rot);
                else error = pi_fallback_axis(u);
                if (unit_try_new_and_get(error, P.allowed_ang, &dir, &depth))
                    solve_generic(b1, b2, a1, a2, true, dir, -depth, P);
                break;
            }
            case 2: {  // cancel_relative_rotation (helper.rs:553-608)
                Vec3 error = quat_scaled_axis(quat_mul(pos2.r, quat_conj(pos1.r)));
                if (unit_try_new_and_get(error, P.allowed_ang, &dir, &depth))
                    solve_generic(b1, b2, a1, a2, true, dir, -depth, P);
                break;
            }
            case 3: {  // project_anchor_to_axis (helper.rs:858-915)
                Vec3 dpt = a2 - a1;
                Vec3 proj = a1 + u * dot3(u, dpt);
                Vec3 error = a2 - proj;
                if (unit_try_new_and_get(error, P.allowed_lin, &dir, &depth))
                    solve_generic(b1, b2, a1, a2, false, dir, -depth, P);
                break;
            }
            case 4: {  // build_linear_limits_position_constraint (unit_constraint.rs:127-197)
                float offset = dot3(u, a2 - a1);
                float error = 0.f;
                dir = u;
                if (j.flags & NB2_JOINT_FLAG_MIN_OFFSET) {
                    error = j.min_offset - offset;
                    dir = -u;
                }
                if (error < 0.f && (j.flags & NB2_JOINT_FLAG_MAX_OFFSET)) {
                    error = offset - j.max_offset;
                    dir = u;
                }
                if (error > P.allowed_lin) solve_generic(b1, b2, a1, a2, false, dir, -error, P);
                break;
            }
            case 5: {  // restore_angle_between_axis (helper.rs:921-998)
                Vec3 sep;
                Quat rot;
                if (quat_rotation_between_axis(u, v, &rot)) sep = quat_scaled_axis(rot);
                else sep = pi_fallback_axis(u);
                float curr;
                if (unit_try_new_and_get(sep, NB2_F32_EPS, &dir, &curr)) {
                    float error = curr - j.angle;
                    if (error < 0.f) {
                        error = -error;
                        dir = -dir;
                    }
                    if (!(error < P.allowed_ang)) solve_generic(b1, b2, a1, a2, true, dir, -error, P);
                }
                break;
            }
            case 6: {  // cancel_relative_translation_wrt_axis (helper.rs:298-359)
                depth = dot3(u, a2 - a1);
                dir = u;
                if (depth < 0.f) {
                    depth = -depth;
                    dir = -u;
                }
                if (depth > P.allowed_lin) solve_generic(b1, b2, a1, a2, false, dir, -depth, P);
                break;
            }
            default: break;
        }
    }
}

struct ContactEval {
    Vec3 world1, world2, normal;
    float depth;
};
// ncollide ContactKinematic::contact
__device__ __forceinline__ bool kinematic_contact(float4 l1, float4 l2, float4 d1, float4 d2, float4 n1,
                                                  const Pose& m1, const Pose& m2, ContactEval* o) {
    const int g1 = __float_as_int(d1.w), g2 = __float_as_int(d2.w);
    Vec3 world1 = pose_point(m1, f4_xyz(l1));
    Vec3 world2 = pose_point(m2, f4_xyz(l2));
    Vec3 normal;
    float depth;
    if (g1 == NB2_GEOM_PLANE && g2 == NB2_GEOM_POINT) {
        normal = quat_rotate(m1.r, f4_xyz(d1));
        depth = -dot3(normal, world2 - world1);
        world1 = world2 + normal * depth;
    } else if (g1 == NB2_GEOM_POINT && g2 == NB2_GEOM_PLANE) {
        Vec3 wn2 = quat_rotate(m2.r, f4_xyz(d2));
        depth = -dot3(wn2, world1 - world2);
        world2 = world1 + wn2 * depth;
        normal = -wn2;
    } else if (g1 != NB2_GEOM_PLANE && g2 != NB2_GEOM_PLANE && (unsigned int)g1 <= NB2_GEOM_PLANE &&
               (unsigned int)g2 <= NB2_GEOM_PLANE) {
        // Point/Point, Line/Line, Line/Point, Point/Line: a Line side is reduced to its point closest to the
        // other side, then the pair is resolved as Point/Point (separated branch: the shapes, which ncollide
        // asks for the tangent cone of the feature, are not part of the contact record; see oracle.cpp)
        if (g1 == NB2_GEOM_LINE && g2 == NB2_GEOM_LINE) {
            const Vec3 e1 = quat_rotate(m1.r, f4_xyz(d1)), e2 = quat_rotate(m2.r, f4_xyz(d2));
            const Vec3 r = world1 - world2;
            const float a = dot3(e1, e1), b = dot3(e1, e2), cc = dot3(e2, e2), d = dot3(e1, r), e = dot3(e2, r);
            const float denom = a * cc - b * b;
            float s1, s2;
            if (denom <= NB2_F32_EPS * a * cc) {
                s1 = 0.f;
                s2 = cc != 0.f ? e / cc : 0.f;
            } else {
                s1 = (b * e - cc * d) / denom;
                s2 = (a * e - b * d) / denom;
            }
            world1 = world1 + e1 * s1;
            world2 = world2 + e2 * s2;
        } else if (g1 == NB2_GEOM_LINE) {
            const Vec3 e1 = quat_rotate(m1.r, f4_xyz(d1));
            const float a = dot3(e1, e1);
            if (a != 0.f) world1 = world1 + e1 * (dot3(e1, world2 - world1) / a);
        } else if (g2 == NB2_GEOM_LINE) {
            const Vec3 e2 = quat_rotate(m2.r, f4_xyz(d2));
            const float a = dot3(e2, e2);
            if (a != 0.f) world2 = world2 + e2 * (dot3(e2, world1 - world2) / a);
        }
        Vec3 n;
        float d;
        if (unit_try_new_and_get(world2 - world1, NB2_F32_EPS, &n, &d)) {
            depth = -d;
            normal = n;
        } else {
            depth = 0.f;
            normal = quat_rotate(m1.r, f4_xyz(n1));
        }
    } else {
        return false;  // Plane/Plane, Plane/Line, Line/Plane: ContactKinematic::contact returns None
    }
    world1 = world1 + normal * l1.w;
    world2 = world2 + normal * (-l2.w);
    depth += l1.w + l2.w;
    o->world1 = world1;
    o->world2 = world2;
    o->normal = normal;
    o->depth = depth;
    return true;
}
__device__ __forceinline__ Pose load_coll(const float* p) {
    Pose r;
    r.t = mk3(p[0], p[1], p[2]);
    r.r = mkq(p[3], p[4], p[5], p[6]);
    return r;
}
}  // namespace nb2
