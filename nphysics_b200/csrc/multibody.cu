// Reduced-coordinate multibodies on device.
//
// Replaces, for the dynamic multibodies of a world: Multibody::update_kinematics / update_body_jacobians /
// update_dynamics / update_inertias / update_acceleration (src/object/multibody.rs:830-863, 404-439, 348-402, 441-628,
// 271-346), Multibody::fill_constraint_geometry (:971-1025) as used by SignoriniCoulombPyramidModel::constraints and
// NonlinearSORProx::update_contact_constraint, the internal constraints of unit joints -- limits and motors --
// (:1027-1152, src/joint/unit_joint.rs), Body::integrate / apply_displacement (:807-824) and the joints
// src/joint/{free,ball,revolute,prismatic,fixed}_joint.rs.
//
// Execution model.  A multibody's rows all act on its own ndofs generalized velocities; rows against static or
// kinematic bodies (and rows between two links of the same multibody) touch nothing else, rows between two multibodies
// touch those two.  The multibodies joined by manifolds form COMPONENTS (labelled on the device every step), and one
// warp runs a component's whole solve in the reference's order -- friction rows of its contacts (two-link class, then
// ground class, each in manifold order), the members' unit-joint rows, normal rows per sweep (sor_prox.rs:159-179);
// internal position constraints, then its contacts per position iteration (nonlinear_sor_prox.rs:33-54).  A multibody
// that only touches the ground is a component of its own (mj_lambda in registers, lanes over its dofs); the batch
// dimension is the number of components (10 000 ragdolls = 10 000 warps).  The multibody path runs beside the
// rigid-body path on the same stream and shares its manifold / contact records, body poses (a link is a
// NB2_BODY_MULTIBODY_LINK body record whose pose the kinematics write) and the per-contact impulse cache.
// Rows between a multibody link and a DYNAMIC rigid body would couple the two paths; they are detected and reported
// (NB2_ERR_UNSUPPORTED through the validation flags), not solved.
//
// Arithmetic follows oracle/multibody.inc expression by expression (-fmad=false): the parity tests compare
// coordinates, velocities and impulses at 1e-5.
#include "solve_position.cuh"

#include <vector>

namespace nb2 {

static const int MB_TPB = 64;
#define NB2_MB_CACHE_WINDOW 64  // how far from its previous index a contact's cached impulse is looked for
#define NB2_MB_MANIFOLD_CAP 32  // manifolds one multibody can be in contact through (more: reported, ignored)

// per multibody (built on the host at upload)
struct MbMeta {
    uint32_t first_link, n_links, ndofs, flags;
    uint32_t dof_off;   // offset of its dofs in the dof-indexed arrays
    uint32_t jac_off;   // floats: n_links blocks of 6 x ndofs (column-major): body jacobians; same offsets in the Coriolis pool
    uint32_t mass_off;  // floats: ndofs x ndofs LU block
    uint32_t has_internal;
};
// per link: description, joint state and derived state
struct MbLinkDev {
    int parent, type;
    uint32_t flags;
    int body;
    uint32_t assembly, ndofs;
    float parent_shift[3], body_shift[3], axis[3];
    float local_com[3], mass, local_inertia[9];
    float min_pos, max_pos, motor_velocity, motor_max_velocity, motor_max_force;
    float free_t[3], free_q[4];  // FreeJoint.position / FixedJoint.body_to_parent
    float rot[4];                // BallJoint.rot / RevoluteJoint.rot
    float coord;                 // angle / offset
    float impulses[3];
    // derived (update_kinematics / update_dynamics)
    float l2w_t[3], l2w_q[4], p2w_q[4], com[3];
    float vel[6], vwj[6], vdwj[6], inertia[9];
    float jc[18];  // ball: jacobian_v (9, row-major), jacobian_dot_v (9); revolute: jacobian (6), jacobian_dot.lin (3), jacobian_dot_veldiff.lin (3)
};
// one velocity row of a multibody: J and M^-1 J (ndofs floats each) live in the row pool at 2 * nd_stride * row
struct MbRow {
    float rhs, r, imp, lim;  // lim: friction coefficient (Dependent) | max force (motor)
    int kind;                // NB2_ROW_*
    int dep;                 // Dependent: row index (within the multibody) of the contact's normal row
    uint32_t contact;        // contact index, or 0xFFFFFFFF for an internal row
    int slot;                // contact rows: 0 normal, 1 / 2 tangents; internal rows: link index * 3 + (0 motor, 1 min, 2 max)
    int other;               // contact rows between two multibodies: the other one (its J / M^-1 J follow this one's in the pool), else -1
    int two_sided;           // both parts are multibody links (the reference's `unilateral` / `bilateral` classes, solved before the ground classes)
};

struct MbView {
    const MbMeta* meta;
    MbLinkDev* links;
    float *vel, *damp, *acc, *ext, *lam;  // [total dofs]
    float *jac, *cor;                     // body jacobians, Coriolis matrices
    float* icd;                           // 6 x ndofs scratch per multibody at 6 * dof_off
    float* mass;                          // LU blocks
    float* accw;                          // workspace.accs: 6 floats per link
    int* piv;                             // [total dofs]
    uint32_t n_mb;
};

__device__ __forceinline__ Vec3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, Vec3 v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}
__device__ __forceinline__ Quat ldq(const float* p) { return mkq(p[0], p[1], p[2], p[3]); }
__device__ __forceinline__ void stq(float* p, Quat q) {
    p[0] = q.i;
    p[1] = q.j;
    p[2] = q.k;
    p[3] = q.w;
}
__device__ __forceinline__ Mat3 ldm(const float* p) {
    Mat3 m;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) m.m[r][c] = p[r * 3 + c];
    return m;
}
__device__ __forceinline__ void stm(float* p, const Mat3& m) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) p[r * 3 + c] = m.m[r][c];
}
// nalgebra's dot: 8 interleaved accumulators from 8 rows on (oracle/multibody.inc::dotn)
__device__ float mb_dot(int n, const float* a, const float* b) {
    float res = 0.f;
    if (n < 8) {
        for (int k = 0; k < n; ++k) res += a[k] * b[k];
        return res;
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int i = 0;
    while (n - i >= 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += a[i + k] * b[i + k];
        i += 8;
    }
    res += acc[0] + acc[4];
    res += acc[1] + acc[5];
    res += acc[2] + acc[6];
    res += acc[3] + acc[7];
    for (; i < n; ++i) res += a[i] * b[i];
    return res;
}
__device__ __forceinline__ float dot6v(const float* a, const float* b) {
    float res = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) res += a[k] * b[k];
    return res;
}
__device__ __forceinline__ Mat3 cross_matrix_tr(Vec3 v) { return mat_transpose(cross_matrix(v)); }

// ------------------------------------------------------------------------------------------------ joints
__device__ Pose mbj_body_to_parent(const MbLinkDev& l) {
    Pose p;
    switch (l.type) {
        case NB2_MBJ_FREE:
            p.t = ld3(l.free_t);
            p.r = ldq(l.free_q);
            return p;
        case NB2_MBJ_BALL:
        case NB2_MBJ_REVOLUTE:
            p.r = ldq(l.rot);
            p.t = ld3(l.parent_shift) - quat_rotate(p.r, ld3(l.body_shift));
            return p;
        case NB2_MBJ_PRISMATIC:
            p.r = mkq(0.f, 0.f, 0.f, 1.f);
            p.t = (ld3(l.parent_shift) - ld3(l.body_shift)) + ld3(l.axis) * l.coord;
            return p;
        default: {
            Pose pt, bt, f;
            pt.t = ld3(l.parent_shift);
            pt.r = mkq(0.f, 0.f, 0.f, 1.f);
            bt.t = ld3(l.body_shift);
            bt.r = pt.r;
            f.t = ld3(l.free_t);
            f.r = ldq(l.free_q);
            return pose_mul(pose_mul(pt, f), bt);
        }
    }
}
__device__ void mbj_update_jacobians(MbLinkDev& l, const float* vels) {
    if (l.type == NB2_MBJ_BALL) {
        Vec3 shift = quat_rotate(ldq(l.rot), -ld3(l.body_shift));
        Vec3 angvel = mk3(vels[0], vels[1], vels[2]);
        stm(l.jc, cross_matrix_tr(shift));
        stm(l.jc + 9, cross_matrix_tr(cross3(angvel, shift)));
    } else if (l.type == NB2_MBJ_REVOLUTE) {
        Vec3 axis = ld3(l.axis);
        Vec3 shift = quat_rotate(ldq(l.rot), -ld3(l.body_shift));
        Vec3 sdv = cross3(axis, shift);
        st3(l.jc, cross3(axis, shift));
        st3(l.jc + 3, axis);
        Vec3 jdv = cross3(axis, sdv);
        st3(l.jc + 9, jdv);
        st3(l.jc + 6, jdv * vels[0]);
    }
}
// out: 6 x ndofs, column-major
__device__ void mbj_jacobian(const MbLinkDev& l, Quat tr, float* out) {
    for (uint32_t k = 0; k < 6 * l.ndofs; ++k) out[k] = 0.f;
    switch (l.type) {
        case NB2_MBJ_FREE:
            for (int k = 0; k < 6; ++k) out[k * 6 + k] = 1.f;
            break;
        case NB2_MBJ_BALL: {
            Mat3 rotmat = quat_to_matrix(tr);
            Mat3 top = mat_mul(rotmat, ldm(l.jc));
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) {
                    out[c * 6 + r] = top.m[r][c];
                    out[c * 6 + 3 + r] = rotmat.m[r][c];
                }
            break;
        }
        case NB2_MBJ_REVOLUTE: {
            st3(out, quat_rotate(tr, ld3(l.jc)));
            st3(out + 3, quat_rotate(tr, ld3(l.jc + 3)));
            break;
        }
        case NB2_MBJ_PRISMATIC: st3(out, quat_rotate(tr, ld3(l.axis))); break;
        default: break;
    }
}
__device__ void mbj_jacobian_dot(const MbLinkDev& l, Quat tr, float* out) {
    for (uint32_t k = 0; k < 6 * l.ndofs; ++k) out[k] = 0.f;
    if (l.type == NB2_MBJ_BALL) {
        Mat3 top = mat_mul(quat_to_matrix(tr), ldm(l.jc + 9));
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) out[c * 6 + r] = top.m[r][c];
    } else if (l.type == NB2_MBJ_REVOLUTE) {
        st3(out, quat_rotate(tr, ld3(l.jc + 6)));
        st3(out + 3, quat_rotate(tr, mk3(0.f, 0.f, 0.f)));
    }
}
__device__ void mbj_jacobian_dot_veldiff(const MbLinkDev& l, Quat tr, const float* acc, float* out) {
    for (uint32_t k = 0; k < 6 * l.ndofs; ++k) out[k] = 0.f;
    if (l.type == NB2_MBJ_BALL) {
        Vec3 angvel = mk3(acc[0], acc[1], acc[2]);
        Mat3 res = mat_mul(mat_mul(quat_to_matrix(tr), cross_matrix(angvel)), ldm(l.jc));
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) out[c * 6 + r] = res.m[r][c];
    } else if (l.type == NB2_MBJ_REVOLUTE) {
        st3(out, quat_rotate(tr, ld3(l.jc + 9)) * acc[0]);
        st3(out + 3, quat_rotate(tr, mk3(0.f, 0.f, 0.f)) * acc[0]);
    }
}
__device__ void mbj_jmul(const MbLinkDev& l, const float* acc, Vec3* lin, Vec3* ang) {
    *lin = mk3(0.f, 0.f, 0.f);
    *ang = *lin;
    switch (l.type) {
        case NB2_MBJ_FREE:
            *lin = mk3(acc[0], acc[1], acc[2]);
            *ang = mk3(acc[3], acc[4], acc[5]);
            break;
        case NB2_MBJ_BALL:
            *ang = mk3(acc[0], acc[1], acc[2]);
            *lin = mat_vec(ldm(l.jc), *ang);
            break;
        case NB2_MBJ_REVOLUTE:
            *lin = ld3(l.jc) * acc[0];
            *ang = ld3(l.jc + 3) * acc[0];
            break;
        case NB2_MBJ_PRISMATIC: *lin = ld3(l.axis) * acc[0]; break;
        default: break;
    }
}
__device__ void mbj_jdotmul(const MbLinkDev& l, const float* acc, Vec3* lin, Vec3* ang) {
    *lin = mk3(0.f, 0.f, 0.f);
    *ang = *lin;
    if (l.type == NB2_MBJ_BALL) *lin = mat_vec(ldm(l.jc + 9), mk3(acc[0], acc[1], acc[2]));
    else if (l.type == NB2_MBJ_REVOLUTE) {
        *lin = ld3(l.jc + 6) * acc[0];
        *ang = mk3(0.f, 0.f, 0.f) * acc[0];
    }
}
__device__ void mbj_free_disp(MbLinkDev& l, Vec3 lin, Vec3 ang) {
    Quat dr = quat_from_scaled_axis(ang);
    st3(l.free_t, lin + ld3(l.free_t));
    stq(l.free_q, quat_mul(dr, ldq(l.free_q)));
}
__device__ void mbj_integrate(MbLinkDev& l, float dt, const float* v) {
    switch (l.type) {
        case NB2_MBJ_FREE: mbj_free_disp(l, mk3(v[0], v[1], v[2]) * dt, mk3(v[3], v[4], v[5]) * dt); break;
        case NB2_MBJ_BALL: {
            Vec3 aa = mk3(v[0], v[1], v[2]) * dt, ax;
            float angle;
            Quat disp = unit_try_new_and_get(aa, 0.f, &ax, &angle) ? quat_from_axis_angle(ax, angle) : mkq(0.f, 0.f, 0.f, 1.f);
            stq(l.rot, quat_mul(disp, ldq(l.rot)));
            break;
        }
        case NB2_MBJ_REVOLUTE:
            l.coord += v[0] * dt;
            stq(l.rot, quat_from_axis_angle(ld3(l.axis), l.coord));
            break;
        case NB2_MBJ_PRISMATIC: l.coord += v[0] * dt; break;
        default: break;
    }
}
__device__ void mbj_apply_displacement(MbLinkDev& l, const float* d) {
    switch (l.type) {
        case NB2_MBJ_FREE: mbj_free_disp(l, mk3(d[0], d[1], d[2]), mk3(d[3], d[4], d[5])); break;
        case NB2_MBJ_BALL: stq(l.rot, quat_mul(quat_from_scaled_axis(mk3(d[0], d[1], d[2])), ldq(l.rot))); break;
        case NB2_MBJ_REVOLUTE:
            l.coord += d[0];
            stq(l.rot, quat_from_axis_angle(ld3(l.axis), l.coord));
            break;
        case NB2_MBJ_PRISMATIC: l.coord += d[0]; break;
        default: break;
    }
}

// ------------------------------------------------------------------------------------------------ multibody
struct Proxies {  // the links' body records as the rest of the library sees them
    PoseQuads pos_t, pos_q;
    float4* com_im;
    float4* vel;
};
__device__ __forceinline__ float& J_at(float* jac, uint32_t nd, uint32_t link, uint32_t r, uint32_t c) {
    return jac[(size_t)link * 6 * nd + (size_t)c * 6 + r];
}

// ---- lane groups.  Every routine below runs either on one thread (G = 1: the position solve, where a multibody's
// kinematics are redone after every displacement by the thread that owns it) or on a warp (G = 32: one warp per
// multibody).  The n_links x ndofs matrices are split by COLUMN: lane c % G owns column c of every link's body
// jacobian and Coriolis matrix and column c of the mass matrix, so the recurrences from parent to child stay inside a
// lane and only the products that mix columns (J^T M J, the LU) need a warp barrier.  Per-link scalars (poses,
// velocities, inertias) are computed by all lanes alike.  Each element is produced by the same expression as on one
// thread, so both shapes give the same bits.
template <int G>
__device__ __forceinline__ void mb_sync() {
    if (G > 1) __syncwarp();
}

// update_kinematics (:830-863) + update_body_jacobians (:404-439)
template <int G>
__device__ void mb_update_kinematics(const MbView& V, const MbMeta& M, const Proxies& P, uint32_t lane) {
    MbLinkDev* L = V.links + M.first_link;
    const uint32_t nd = M.ndofs;
    float* jac = V.jac + M.jac_off;
    const float* vel = V.vel + M.dof_off;
    for (uint32_t i = 0; i < M.n_links; ++i) {
        MbLinkDev& rb = L[i];
        mbj_update_jacobians(rb, vel + rb.assembly);
        Pose l2p = mbj_body_to_parent(rb), l2w;
        Quat p2w = mkq(0.f, 0.f, 0.f, 1.f);
        if (i == 0) {
            l2w = l2p;
        } else {
            const MbLinkDev& pr = L[rb.parent];
            Pose pw;
            pw.t = ld3(pr.l2w_t);
            pw.r = ldq(pr.l2w_q);
            l2w = pose_mul(pw, l2p);
            p2w = pw.r;
        }
        st3(rb.l2w_t, l2w.t);
        stq(rb.l2w_q, l2w.r);
        stq(rb.p2w_q, p2w);
        Vec3 com = pose_point(l2w, ld3(rb.local_com));
        st3(rb.com, com);
        if (rb.body >= 0 && lane == 0) {
            P.pos_t[rb.body] = xyz_f4(l2w.t, 0.f);
            P.pos_q[rb.body] = quat_f4(l2w.r);
            P.com_im[rb.body] = xyz_f4(com, 0.f);
        }
        mb_sync<G>();
    }
    float jj[36];
    for (uint32_t i = 0; i < M.n_links; ++i) {
        const MbLinkDev& rb = L[i];
        if (i != 0) {
            const MbLinkDev& pr = L[rb.parent];
            Mat3 shift_tr = cross_matrix_tr(ld3(rb.com) - ld3(pr.com));
            for (uint32_t c = lane; c < nd; c += G) {
                float pj[6];
                for (int r = 0; r < 6; ++r) pj[r] = J_at(jac, nd, rb.parent, r, c);
                for (int r = 0; r < 3; ++r) {
                    float s = 0.f;
                    for (int k = 0; k < 3; ++k) s += shift_tr.m[r][k] * pj[3 + k];
                    J_at(jac, nd, i, r, c) = s + pj[r];
                }
                for (int r = 3; r < 6; ++r) J_at(jac, nd, i, r, c) = pj[r];
            }
        } else {
            for (uint32_t c = lane; c < nd; c += G)
                for (int r = 0; r < 6; ++r) J_at(jac, nd, i, r, c) = 0.f;
        }
        mbj_jacobian(rb, ldq(rb.p2w_q), jj);
        for (uint32_t c = 0; c < rb.ndofs; ++c)
            if ((rb.assembly + c) % G == lane)
                for (int r = 0; r < 6; ++r) J_at(jac, nd, i, r, rb.assembly + c) += jj[c * 6 + r];
    }
    mb_sync<G>();
}

// nalgebra LU::new (partial pivoting) on the nd x nd column-major block m; piv[i] = row swapped with row i
template <int G>
__device__ void mb_lu_factor(float* m, int* piv, int n, uint32_t lane) {
    for (int i = 0; i < n; ++i) {
        int p = i;
        float best = fabsf(m[i * n + i]);
        for (int r = i + 1; r < n; ++r)
            if (fabsf(m[i * n + r]) > best) {
                best = fabsf(m[i * n + r]);
                p = r;
            }
        if (lane == 0) piv[i] = p;
        const float diag = m[i * n + p];
        if (diag == 0.f) continue;
        mb_sync<G>();  // everybody has read column i before rows are swapped
        if (p != i)
            for (int c = (int)lane; c < n; c += G) {
                float t = m[c * n + i];
                m[c * n + i] = m[c * n + p];
                m[c * n + p] = t;
            }
        mb_sync<G>();
        const float inv_diag = 1.f / diag;
        for (int c = i + 1 + (int)((lane + G - (uint32_t)((i + 1) % G)) % G); c < n; c += G) {
            const float pivot_row = m[c * n + i];
            for (int r = i + 1; r < n; ++r) m[c * n + r] = (-pivot_row) * (m[i * n + r] * inv_diag) + m[c * n + r];
        }
        mb_sync<G>();
        if ((uint32_t)(i % G) == lane)
            for (int r = i + 1; r < n; ++r) m[i * n + r] *= inv_diag;
        mb_sync<G>();
    }
}
// LU::solve_mut on b (memory every lane of the group sees)
template <int G>
__device__ bool mb_lu_solve(const float* m, const int* piv, int n, float* b, uint32_t lane) {
    if (lane == 0)
        for (int i = 0; i < n; ++i)
            if (piv[i] != i) {
                float t = b[i];
                b[i] = b[piv[i]];
                b[piv[i]] = t;
            }
    mb_sync<G>();
    for (int i = 0; i + 1 < n; ++i) {
        const float coeff = b[i];
        for (int r = i + 1 + (int)lane; r < n; r += G) b[r] = (-coeff) * m[i * n + r] + b[r];
        mb_sync<G>();
    }
    bool ok = true;
    for (int i = n - 1; i >= 0; --i) {
        const float diag = m[i * n + i];
        if (diag == 0.f) {
            ok = false;
            break;
        }
        if (lane == 0) b[i] /= diag;
        mb_sync<G>();
        const float coeff = b[i];
        for (int r = (int)lane; r < i; r += G) b[r] = (-coeff) * m[i * n + r] + b[r];
        mb_sync<G>();
    }
    return ok;
}
// the same on a vector private to the calling thread
__device__ __forceinline__ bool mb_lu_solve(const float* m, const int* piv, int n, float* b) { return mb_lu_solve<1>(m, piv, n, b, 0u); }

// update_dynamics (:348-402) + update_inertias (:441-628)
template <int G>
__device__ void mb_update_dynamics(const MbView& V, const MbMeta& M, const Proxies& P, float dt, uint32_t lane) {
    MbLinkDev* L = V.links + M.first_link;
    const uint32_t nd = M.ndofs;
    float* jac = V.jac + M.jac_off;
    float* cor = V.cor + M.jac_off;  // per link: rows 0..2 coriolis_v, rows 3..5 coriolis_w
    float* icd = V.icd + (size_t)6 * M.dof_off;
    float* mass = V.mass + M.mass_off;
    const float* vel = V.vel + M.dof_off;
    const float* damp = V.damp + M.dof_off;
    for (uint32_t i = 0; i < M.n_links; ++i) {
        MbLinkDev& rb = L[i];
        Vec3 wl, wa, dl, da;
        mbj_jmul(rb, vel + rb.assembly, &wl, &wa);
        mbj_jdotmul(rb, vel + rb.assembly, &dl, &da);
        Vec3 vl, va;
        if (i == 0) {
            vl = wl;
            va = wa;
        } else {
            const MbLinkDev& pr = L[rb.parent];
            Quat pq = ldq(pr.l2w_q);
            dl = quat_rotate(pq, dl);
            da = quat_rotate(pq, da);
            wl = quat_rotate(pq, wl);
            wa = quat_rotate(pq, wa);
            vl = ld3(pr.vel) + wl;
            va = ld3(pr.vel + 3) + wa;
            Vec3 shift = ld3(rb.com) - ld3(pr.com);
            vl = vl + cross3(ld3(pr.vel + 3), shift);
        }
        st3(rb.vdwj, dl);
        st3(rb.vdwj + 3, da);
        st3(rb.vwj, wl);
        st3(rb.vwj + 3, wa);
        st3(rb.vel, vl);
        st3(rb.vel + 3, va);
        if (rb.body >= 0 && lane == 0) {
            P.vel[2 * rb.body] = xyz_f4(vl, 0.f);
            P.vel[2 * rb.body + 1] = xyz_f4(va, 0.f);
        }
        Mat3 rot = quat_to_matrix(ldq(rb.l2w_q));
        stm(rb.inertia, mat_mul(mat_mul(rot, ldm(rb.local_inertia)), mat_transpose(rot)));
        mb_sync<G>();
    }
    for (uint32_t c = lane; c < nd; c += G)
        for (uint32_t r = 0; r < nd; ++r) mass[c * nd + r] = 0.f;
    float work[6], t1[36], t2[36];
    for (uint32_t i = 0; i < M.n_links; ++i) {
        const MbLinkDev& rb = L[i];
        const Mat3 ang_inertia = ldm(rb.inertia);
        const Vec3 w = ld3(rb.vel + 3);
        Mat3 aug = ang_inertia;
        {
            Mat3 a = mat_mul(cross_matrix(w), ang_inertia), b = cross_matrix(mat_vec(ang_inertia, w));
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) aug.m[r][c] = aug.m[r][c] + (a.m[r][c] - b.m[r][c]) * dt;
        }
        for (uint32_t j = lane; j < nd; j += G) {  // quadform: column j of J^T M J
            float cj[6];
            for (int r = 0; r < 6; ++r) cj[r] = J_at(jac, nd, i, r, j);
            for (int r = 0; r < 3; ++r) work[r] = rb.mass * cj[r];
            Vec3 wa = mat_vec(aug, mk3(cj[3], cj[4], cj[5]));
            work[3] = wa.x;
            work[4] = wa.y;
            work[5] = wa.z;
            for (uint32_t r = 0; r < nd; ++r) mass[j * nd + r] = dot6v(&J_at(jac, nd, i, 0, r), work) + mass[j * nd + r];
        }
        // Coriolis matrix
        if (i != 0) {
            const MbLinkDev& pr = L[rb.parent];
            const Mat3 parent_w = cross_matrix(ld3(pr.vel + 3));
            const Mat3 shift_tr = cross_matrix_tr(ld3(rb.com) - ld3(pr.com));
            const Mat3 dvel_tr = cross_matrix_tr(ld3(rb.vel) - ld3(pr.vel));
            const Mat3 vwj_tr = cross_matrix_tr(ld3(rb.vwj));
            const Mat3 vwj_w = cross_matrix(ld3(rb.vwj + 3));
            for (uint32_t c = lane; c < nd; c += G) {
                Vec3 pjv = mk3(J_at(jac, nd, rb.parent, 0, c), J_at(jac, nd, rb.parent, 1, c), J_at(jac, nd, rb.parent, 2, c));
                Vec3 pjw = mk3(J_at(jac, nd, rb.parent, 3, c), J_at(jac, nd, rb.parent, 4, c), J_at(jac, nd, rb.parent, 5, c));
                Vec3 rjv = mk3(J_at(jac, nd, i, 0, c), J_at(jac, nd, i, 1, c), J_at(jac, nd, i, 2, c));
                Vec3 v = mk3(J_at(cor, nd, rb.parent, 0, c), J_at(cor, nd, rb.parent, 1, c), J_at(cor, nd, rb.parent, 2, c));
                Vec3 pw = mk3(J_at(cor, nd, rb.parent, 3, c), J_at(cor, nd, rb.parent, 4, c), J_at(cor, nd, rb.parent, 5, c));
                v = mat_vec(shift_tr, pw) + v;
                v = mat_vec(dvel_tr, pjw) + v;
                v = mat_vec(vwj_tr, pjw) + v;
                v = mat_vec(parent_w, rjv) + v;
                v = mat_vec(parent_w, pjv) * -1.f + v;
                Vec3 ww = mat_vec(vwj_w, pjw) * -1.f + pw;
                J_at(cor, nd, i, 0, c) = v.x;
                J_at(cor, nd, i, 1, c) = v.y;
                J_at(cor, nd, i, 2, c) = v.z;
                J_at(cor, nd, i, 3, c) = ww.x;
                J_at(cor, nd, i, 4, c) = ww.y;
                J_at(cor, nd, i, 5, c) = ww.z;
            }
            mbj_jacobian(rb, ldq(pr.l2w_q), t1);
            for (uint32_t c = 0; c < rb.ndofs; ++c) {
                const uint32_t cc = rb.assembly + c;
                if (cc % G != lane) continue;
                Vec3 a = mat_vec(parent_w, mk3(t1[c * 6], t1[c * 6 + 1], t1[c * 6 + 2]));
                Vec3 b = mat_vec(parent_w, mk3(t1[c * 6 + 3], t1[c * 6 + 4], t1[c * 6 + 5]));
                J_at(cor, nd, i, 0, cc) += a.x;
                J_at(cor, nd, i, 1, cc) += a.y;
                J_at(cor, nd, i, 2, cc) += a.z;
                J_at(cor, nd, i, 3, cc) += b.x;
                J_at(cor, nd, i, 4, cc) += b.y;
                J_at(cor, nd, i, 5, cc) += b.z;
            }
        } else {
            for (uint32_t c = lane; c < nd; c += G)
                for (int r = 0; r < 6; ++r) J_at(cor, nd, i, r, c) = 0.f;
        }
        mbj_jacobian_dot(rb, ldq(rb.p2w_q), t1);
        mbj_jacobian_dot_veldiff(rb, ldq(rb.p2w_q), vel + rb.assembly, t2);
        for (uint32_t c = 0; c < rb.ndofs; ++c) {
            const uint32_t cc = rb.assembly + c;
            if (cc % G != lane) continue;
            for (int r = 0; r < 6; ++r) J_at(cor, nd, i, r, cc) += t1[c * 6 + r];
            for (int r = 0; r < 6; ++r) J_at(cor, nd, i, r, cc) += t2[c * 6 + r];
        }
        for (uint32_t c = lane; c < nd; c += G) {
            for (int r = 0; r < 3; ++r) icd[c * 6 + r] = J_at(cor, nd, i, r, c) * (rb.mass * dt);
            Vec3 wv = mat_vec(ang_inertia, mk3(J_at(cor, nd, i, 3, c), J_at(cor, nd, i, 4, c), J_at(cor, nd, i, 5, c)));
            icd[c * 6 + 3] = dt * wv.x;
            icd[c * 6 + 4] = dt * wv.y;
            icd[c * 6 + 5] = dt * wv.z;
            for (uint32_t r = 0; r < nd; ++r) mass[c * nd + r] = dot6v(&J_at(jac, nd, i, 0, r), icd + c * 6) + mass[c * nd + r];
        }
    }
    for (uint32_t k = lane; k < nd; k += G) mass[k * nd + k] += damp[k] * dt;
    mb_sync<G>();
    mb_lu_factor<G>(mass, V.piv + M.dof_off, (int)nd, lane);
}

// update_acceleration (:271-346); ext = dt * acceleration, mj_lambda = 0
template <int G>
__device__ void mb_update_acceleration(const MbView& V, const MbMeta& M, Vec3 gravity, float dt, uint32_t lane) {
    MbLinkDev* L = V.links + M.first_link;
    const uint32_t nd = M.ndofs;
    const float* jac = V.jac + M.jac_off;
    float* acc = V.acc + M.dof_off;
    const float* vel = V.vel + M.dof_off;
    const float* damp = V.damp + M.dof_off;
    float* accs = V.accw + (size_t)6 * M.first_link;  // workspace.accs
    for (uint32_t c = lane; c < nd; c += G) acc[c] = 0.f;
    for (uint32_t i = 0; i < M.n_links; ++i) {
        const MbLinkDev& rb = L[i];
        Vec3 al = ld3(rb.vdwj), aa = ld3(rb.vdwj + 3);
        if (i != 0) {
            const MbLinkDev& pr = L[rb.parent];
            const Vec3 pal = ld3(accs + 6 * rb.parent), paa = ld3(accs + 6 * rb.parent + 3);
            const Vec3 pw = ld3(pr.vel + 3);
            al = al + pal;
            aa = aa + paa;
            al = al + cross3(pw, ld3(rb.vwj));
            aa = aa + cross3(pw, ld3(rb.vwj + 3));
            Vec3 shift = ld3(rb.com) - ld3(pr.com);
            Vec3 dvel = ld3(rb.vel) - ld3(pr.vel);
            al = al + cross3(pw, dvel);
            al = al + cross3(paa, shift);
        }
        st3(accs + 6 * i, al);
        st3(accs + 6 * i + 3, aa);
        mb_sync<G>();
        const Mat3 inertia = ldm(rb.inertia);
        const Vec3 w = ld3(rb.vel + 3);
        Vec3 gf = (M.flags & NB2_BODY_FLAG_GRAVITY) ? gravity * rb.mass : mk3(0.f, 0.f, 0.f);
        Vec3 gyro = cross3(w, mat_vec(inertia, w));
        Vec3 fl = gf - al * rb.mass;
        Vec3 fa = (-gyro) - mat_vec(inertia, aa);
        const float f[6] = {fl.x, fl.y, fl.z, fa.x, fa.y, fa.z};
        for (uint32_t c = lane; c < nd; c += G) acc[c] = dot6v(&jac[(size_t)i * 6 * nd + (size_t)c * 6], f) + acc[c];
    }
    for (uint32_t c = lane; c < nd; c += G) {
        float a = 0.f + acc[c];  // + generalized forces (none through this ABI)
        acc[c] = -1.f * damp[c] * vel[c] + a;
    }
    mb_sync<G>();
    mb_lu_solve<G>(V.mass + M.mass_off, V.piv + M.dof_off, (int)nd, acc, lane);
    float* ext = V.ext + M.dof_off;
    float* lam = V.lam + M.dof_off;
    for (uint32_t c = lane; c < nd; c += G) {
        ext[c] = dt * acc[c];
        lam[c] = 0.f;
    }
}

// mode 0: kinematics + dynamics (upload); 1: + accelerations (start of a step); 2: kinematics + link velocities only
// (end of a step: the next step recomputes the mass matrix from the same state, mechanical_world.rs:343-346 / :230-243).
// One warp per multibody.  `stage_words` != 0: the block's dynamic shared memory holds one scratch region per warp
// (body jacobians, Coriolis matrices, mass matrix, pivots); the refresh then runs out of shared memory and only the
// jacobians and the LU go back to global memory, once, coalesced.
__device__ void mb_link_velocities(const MbView& V, const MbMeta& M, const Proxies& P, uint32_t lane) {
    MbLinkDev* L = V.links + M.first_link;
    const float* vel = V.vel + M.dof_off;
    for (uint32_t i = 0; i < M.n_links; ++i) {
        MbLinkDev& rb = L[i];
        Vec3 wl, wa;
        mbj_jmul(rb, vel + rb.assembly, &wl, &wa);
        Vec3 vl = wl, va = wa;
        if (i != 0) {
            const MbLinkDev& pr = L[rb.parent];
            Quat pq = ldq(pr.l2w_q);
            wl = quat_rotate(pq, wl);
            wa = quat_rotate(pq, wa);
            vl = ld3(pr.vel) + wl;
            va = ld3(pr.vel + 3) + wa;
            vl = vl + cross3(ld3(pr.vel + 3), ld3(rb.com) - ld3(pr.com));
        }
        st3(rb.vel, vl);
        st3(rb.vel + 3, va);
        if (rb.body >= 0 && lane == 0) {
            P.vel[2 * rb.body] = xyz_f4(vl, 0.f);
            P.vel[2 * rb.body + 1] = xyz_f4(va, 0.f);
        }
        __syncwarp();
    }
}
#define MB_WPB 4  // warps (multibodies) per block of the warp-per-multibody kernels
__global__ void __launch_bounds__(32 * MB_WPB) k_mb_refresh(MbView V, Proxies P, float dt, Vec3 gravity, int mode, uint32_t stage_words,
                                                            uint32_t jac_words) {
    extern __shared__ float mb_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= V.n_mb) return;
    MbMeta M = V.meta[m];
    if (mode == 2) {
        mb_update_kinematics<32>(V, M, P, lane);
        mb_link_velocities(V, M, P, lane);
        return;
    }
    if (stage_words == 0) {
        mb_update_kinematics<32>(V, M, P, lane);
        mb_update_dynamics<32>(V, M, P, dt, lane);
        if (mode == 1) mb_update_acceleration<32>(V, M, gravity, dt, lane);
        return;
    }
    // per-warp region: [jac | cor | mass | icd | piv]
    const uint32_t nd = M.ndofs, jw = M.n_links * 6 * nd;
    float* reg = mb_smem + (size_t)warp * stage_words;
    MbView W = V;
    MbMeta L = M;
    W.jac = reg;
    W.cor = reg + jac_words;
    W.mass = reg + 2 * jac_words;
    W.icd = W.mass + nd * nd;
    W.piv = reinterpret_cast<int*>(W.icd + 6 * nd);
    W.vel = V.vel + M.dof_off;
    W.damp = V.damp + M.dof_off;
    W.acc = V.acc + M.dof_off;
    W.ext = V.ext + M.dof_off;
    W.lam = V.lam + M.dof_off;
    L.dof_off = 0;
    L.jac_off = 0;
    L.mass_off = 0;
    mb_update_kinematics<32>(W, L, P, lane);
    mb_update_dynamics<32>(W, L, P, dt, lane);
    if (mode == 1) mb_update_acceleration<32>(W, L, gravity, dt, lane);
    __syncwarp();
    float* gj = V.jac + M.jac_off;
    for (uint32_t k = lane; k < jw; k += 32) gj[k] = W.jac[k];
    float* gm = V.mass + M.mass_off;
    for (uint32_t k = lane; k < nd * nd; k += 32) gm[k] = W.mass[k];
    int* gp = V.piv + M.dof_off;
    for (uint32_t k = lane; k < nd; k += 32) gp[k] = W.piv[k];
}

// ------------------------------------------------------------------------------------------------ rows
struct MbContacts {
    const nb2_manifold* manifolds;
    const nb2_contact* contacts;
    uint32_t n_manifolds, n_contacts;
    const int* link_of_body;  // global link index or -1
    const int* mb_of_link;    // multibody of a (global) link
    const int* status;        // effective body status
    ConstPoseQuads pos_t, pos_q;
    const float4* vel;
    const float4* com_im;
    uint32_t* mcount;         // [n_mb] manifolds of each multibody
    uint32_t* mlist;          // [n_mb][NB2_MB_MANIFOLD_CAP]
    uint32_t* edges;          // pairs of multibodies joined by a manifold: [2 * edge], capacity n_manifolds
    uint32_t* n_edges;
    uint32_t* flags;          // validation flags of the context ([0] |= bits)
};
#define NB2_FLAG_MB_COUPLED 0x100u   // a manifold couples a multibody with a dynamic body or another multibody
#define NB2_FLAG_MB_OVERFLOW 0x200u  // more than NB2_MB_MANIFOLD_CAP manifolds on one multibody, or the row pool is full

__global__ void k_mb_collect(MbContacts C, const int* __restrict__ mb_of_link) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= C.n_manifolds) return;
    const nb2_manifold& mf = C.manifolds[m];
    if (mf.num_contacts == 0) return;
    const int l1 = C.link_of_body[mf.body1], l2 = C.link_of_body[mf.body2];
    if (l1 < 0 && l2 < 0) return;
    const int s1 = C.status[mf.body1], s2 = C.status[mf.body2];
    int mb;
    if (l1 >= 0 && l2 >= 0) {
        const int a = mb_of_link[l1], b = mb_of_link[l2];
        mb = a < b ? a : b;  // a manifold between two multibodies is owned by the lower one and joins their components
        if (a != b) {
            const uint32_t e = atomicAdd(C.n_edges, 1u);
            C.edges[2 * e] = (uint32_t)a;
            C.edges[2 * e + 1] = (uint32_t)b;
        }
    } else {
        const int other = l1 >= 0 ? s2 : s1;
        if (other == NB2_BODY_DISABLED) return;  // mechanical_world.rs:287-300
        if (other == NB2_BODY_DYNAMIC) {
            atomicOr(&C.flags[0], NB2_FLAG_MB_COUPLED);
            return;
        }
        mb = mb_of_link[l1 >= 0 ? l1 : l2];
    }
    const uint32_t k = atomicAdd(&C.mcount[mb], 1u);
    if (k < NB2_MB_MANIFOLD_CAP) C.mlist[(size_t)mb * NB2_MB_MANIFOLD_CAP + k] = m;
    else atomicOr(&C.flags[0], NB2_FLAG_MB_OVERFLOW);
}

// Components of the multibodies under "share a manifold": label = smallest member.  One block; labels are relaxed
// over the edges until nothing changes (the edge list is empty for multibodies that only touch the ground).
__global__ void k_mb_components(uint32_t n_mb, const uint32_t* __restrict__ edges, const uint32_t* __restrict__ n_edges, uint32_t* comp,
                                uint32_t* ccount) {
    __shared__ int changed;
    for (uint32_t m = threadIdx.x; m < n_mb; m += blockDim.x) {
        comp[m] = m;
        ccount[m] = 0u;
    }
    if (threadIdx.x == 0) ccount[n_mb] = 0u;
    const uint32_t ne = *n_edges;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < ne; e += blockDim.x) {
            const uint32_t a = comp[edges[2 * e]], b = comp[edges[2 * e + 1]];
            if (a < b) {
                atomicMin(&comp[edges[2 * e + 1]], a);
                changed = 1;
            } else if (b < a) {
                atomicMin(&comp[edges[2 * e]], b);
                changed = 1;
            }
        }
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    for (uint32_t m = threadIdx.x; m < n_mb; m += blockDim.x) atomicAdd(&ccount[comp[m]], 1u);
}
// members of every component, after the scan of ccount into coff (any order: the component's warp sorts its few members)
__global__ void k_mb_members(uint32_t n_mb, const uint32_t* __restrict__ comp, const uint32_t* __restrict__ coff, uint32_t* ccursor,
                             uint32_t* cmem) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mb) return;
    const uint32_t root = comp[m];
    cmem[coff[root] + atomicAdd(&ccursor[root], 1u)] = m;
}

struct MbRows {
    MbRow* rows;        // [row_cap]
    float* jw;          // [row_cap][4 * nd_stride]: J, M^-1 J over the owner's dofs, then over the other multibody's
    float4* cpos;       // per contact row triple: local normal of body 1 (position solve), [row_cap / 3 + ...] indexed by the normal row
    uint32_t* row_off;  // [n_mb + 1] first row of each multibody (exclusive scan of row_cnt)
    uint32_t* row_cnt;  // [n_mb] rows assembled for each multibody: [0] friction, then internal, then normal
    uint32_t* seg;      // [n_mb][4]: friction rows, internal unilateral rows, internal bilateral rows, normal rows
    uint32_t nd_stride, row_cap;
    // components (multibodies joined by manifolds): label, member list, scratch for the ordered contact lists
    const uint32_t *comp, *coff, *cmem_c;
    uint32_t* cmem;
    uint32_t* lslot;    // [n_mb] offset of a member's mj_lambda in its component's shared-memory slice
    uint2* corder;      // ordered contacts of the components: (normal row, first friction row); slices from a bump allocator
    uint2* morder;      // scratch of the same shape: a component's manifolds while they are sorted
    uint32_t* cursor;   // the allocator
};

// Multibody::fill_constraint_geometry (:971-1025): J = body_jacobian^T force, WJ = M^-1 J; accumulates (+=) into J / WJ so
// that the two sides of a self contact add up to one row over the multibody's dofs.
__device__ void mb_fill_geometry(const MbView& V, const MbMeta& M, uint32_t link, Vec3 point, bool angular, Vec3 dir, float* tmp,
                                 float* J, float* WJ, bool accumulate) {
    const MbLinkDev& rb = V.links[M.first_link + link];
    const uint32_t nd = M.ndofs;
    Vec3 pos = point - ld3(rb.com);
    Vec3 fl = angular ? mk3(0.f, 0.f, 0.f) : dir;
    Vec3 fa = angular ? dir : cross3(pos, dir);
    const float f[6] = {fl.x, fl.y, fl.z, fa.x, fa.y, fa.z};
    const float* jac = V.jac + M.jac_off + (size_t)link * 6 * nd;
    for (uint32_t c = 0; c < nd; ++c) tmp[c] = dot6v(jac + (size_t)c * 6, f);
    for (uint32_t c = 0; c < nd; ++c) J[c] = accumulate ? J[c] + tmp[c] : tmp[c];
    mb_lu_solve(V.mass + M.mass_off, V.piv + M.dof_off, (int)nd, tmp);
    for (uint32_t c = 0; c < nd; ++c) WJ[c] = accumulate ? WJ[c] + tmp[c] : tmp[c];
}

struct MbCache {
    const unsigned long long* ckey_prev;
    const float4* imp_prev;
    unsigned int n_prev;
    float4* imp_cur;
    unsigned long long* ckey_cur;
};

// One thread per multibody: the rows of its contacts (SignoriniCoulombPyramidModel::constraints,
// signorini_coulomb_pyramid_model.rs:56-224 with SignoriniModel::build_velocity_constraint, signorini_model.rs:37-138)
// and of its unit joints' motors and limits (unit_joint.rs:39-196).
__global__ void __launch_bounds__(32 * MB_WPB) k_mb_assemble(MbView Vg, MbContacts C, MbRows R, MbCache K, const int* __restrict__ mb_of_link,
                                                             float warmstart_coeff, float restitution_threshold, float inv_dt,
                                                             uint32_t stage_words, uint32_t jac_words, int model) {
    // One warp per multibody; its lanes take the (contact, row) jobs -- every row is one J = J_link^T f and one
    // LU solve, independent of the others -- with the body jacobians and the LU of the mass matrix staged in the warp's
    // slice of shared memory (read by all lanes at the same address: a broadcast).
    extern __shared__ float mb_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= Vg.n_mb) return;
    MbMeta M = Vg.meta[m];
    const uint32_t nd = M.ndofs;
    MbView V = Vg;
    if (stage_words) {
        float* reg = mb_smem + (size_t)warp * stage_words;
        const uint32_t jw = M.n_links * 6 * nd;
        const float* gj = Vg.jac + M.jac_off;
        for (uint32_t k = lane; k < jw; k += 32) reg[k] = gj[k];
        float* sm = reg + jac_words;
        const float* gm = Vg.mass + M.mass_off;
        for (uint32_t k = lane; k < nd * nd; k += 32) sm[k] = gm[k];
        int* sp = reinterpret_cast<int*>(sm + nd * nd);
        const int* gp = Vg.piv + M.dof_off;
        for (uint32_t k = lane; k < nd; k += 32) sp[k] = gp[k];
        V.jac = reg;
        V.mass = sm;
        V.piv = sp;
        V.vel = Vg.vel + M.dof_off;
        V.ext = Vg.ext + M.dof_off;
        M.jac_off = 0;
        M.mass_off = 0;
        M.dof_off = 0;
    }
    uint32_t* list = C.mlist + (size_t)m * NB2_MB_MANIFOLD_CAP;
    const uint32_t nm = min(C.mcount[m], (uint32_t)NB2_MB_MANIFOLD_CAP);
    if (lane == 0)
        for (uint32_t a = 1; a < nm; ++a) {  // manifold order (the atomics filled the list in any order)
            const uint32_t v = list[a];
            uint32_t b = a;
            while (b > 0 && list[b - 1] > v) {
                list[b] = list[b - 1];
                --b;
            }
            list[b] = v;
        }
    __syncwarp();
    const uint32_t base = R.row_off[m];
    const uint32_t cap = R.row_off[m + 1] - base;
    uint32_t ncon = 0;
    for (uint32_t a = 0; a < nm; ++a) ncon += C.manifolds[list[a]].num_contacts;
    // internal rows first counted: motors (bilateral) and active limits (unilateral)
    uint32_t n_uni = 0, n_bil = 0;
    const MbLinkDev* L = V.links + M.first_link;
    for (uint32_t i = 0; i < M.n_links; ++i) {
        const MbLinkDev& l = L[i];
        if (l.type != NB2_MBJ_REVOLUTE && l.type != NB2_MBJ_PRISMATIC) continue;
        if (l.flags & NB2_MBJ_FLAG_MOTOR) ++n_bil;
        if ((l.flags & NB2_MBJ_FLAG_MIN) && l.min_pos - l.coord >= 0.f) ++n_uni;
        if ((l.flags & NB2_MBJ_FLAG_MAX) && -(l.max_pos - l.coord) >= 0.f) ++n_uni;
    }
    if (2 * ncon + n_uni + n_bil + ncon > cap) {  // cannot happen: the offsets were sized from the same counts
        if (lane == 0) atomicOr(&C.flags[0], NB2_FLAG_MB_OVERFLOW);
        ncon = 0;
    }
    if (lane == 0) {
        uint32_t* seg = R.seg + 4 * m;
        seg[0] = 2 * ncon;
        seg[1] = n_uni;
        seg[2] = n_bil;
        seg[3] = ncon;
    }
    const uint32_t fr0 = base, in0 = base + 2 * ncon, no0 = in0 + n_uni + n_bil;
    const size_t rs = (size_t)4 * R.nd_stride;
    const float* vel = V.vel + M.dof_off;
    const float* ext = V.ext + M.dof_off;
    float tmp[NB2_MB_MAX_DOFS];
    // ---- contacts: job = 3 * k + w, k the contact's rank within the multibody (manifold order), w the row
    __syncwarp();
    for (uint32_t job = lane; job < 3 * ncon; job += 32) {
        const uint32_t k = job / 3;
        const int w = (int)(job % 3);
        uint32_t a = 0, before = 0;
        for (; a < nm; ++a) {
            const uint32_t n = C.manifolds[list[a]].num_contacts;
            if (k < before + n) break;
            before += n;
        }
        const nb2_manifold& mf = C.manifolds[list[a]];
        const uint32_t ci = mf.first_contact + (k - before);
        if (ci >= C.n_contacts) continue;
        const int l1 = C.link_of_body[mf.body1], l2 = C.link_of_body[mf.body2];
        const Vec3 surf = mk3(mf.surface_velocity[0], mf.surface_velocity[1], mf.surface_velocity[2]);
        const Quat q1 = f4_quat(C.pos_q[mf.body1]);
        {
            const nb2_contact& c = C.contacts[ci];
            const Vec3 normal = ld3(c.normal), world1 = ld3(c.world1), world2 = ld3(c.world2);
            float4 cached = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c.key != 0ull) {
                // ImpulseCache lookup by ContactId (signorini_coulomb_pyramid_model.rs:104-108): the per-contact arrays of
                // the previous step at the same index (a contact that kept its place), else the nearest indices on
                // both sides -- contacts are dropped and inserted around it, so its old slot is a few entries away.
                // Beyond NB2_MB_CACHE_WINDOW entries the contact starts cold.
                bool hit = ci < K.n_prev && K.ckey_prev[ci] == c.key;
                if (hit) cached = K.imp_prev[ci];
                for (uint32_t d = 1; !hit && d <= NB2_MB_CACHE_WINDOW; ++d) {
                    if (ci >= d && ci - d < K.n_prev && K.ckey_prev[ci - d] == c.key) {
                        cached = K.imp_prev[ci - d];
                        hit = true;
                    } else if (ci + d < K.n_prev && K.ckey_prev[ci + d] == c.key) {
                        cached = K.imp_prev[ci + d];
                        hit = true;
                    }
                }
            }
            const Vec3 center1 = world1 + normal * mf.margin1;
            const Vec3 center2 = world2 - normal * mf.margin2;
            Vec3 t1, t2;
            tangent_basis(normal, &t1, &t2);
            const Vec3 dir = w == 0 ? -normal : (w == 1 ? t1 : t2);
            const float rhs0 = w == 0 ? dot3(normal, surf) : (w == 1 ? dot3(t1, surf) : dot3(t2, surf));
            const uint32_t rowi = w == 0 ? no0 + k : fr0 + 2 * k + (uint32_t)(w - 1);
            {
                float* J = R.jw + (size_t)rowi * rs;
                float* WJ = J + R.nd_stride;
                float out_vel = rhs0;
                bool first = true;
                // a kinematic partner contributes its velocity at the contact point (fill_constraint_geometry's
                // Kinematic branch, rigid_body.rs:693-698); a static one nothing
                auto kinematic_side = [&](int body, Vec3 center, Vec3 d) {
                    if (C.status[body] != NB2_BODY_KINEMATIC) return;
                    const Vec3 pos = center - f4_xyz(C.com_im[body]);
                    const Vec3 fa = cross3(pos, d);
                    const float f[6] = {d.x, d.y, d.z, fa.x, fa.y, fa.z};
                    const float4 vl = C.vel[2 * body], va = C.vel[2 * body + 1];
                    const float v[6] = {vl.x, vl.y, vl.z, va.x, va.y, va.z};
                    out_vel += dot6v(f, v);
                };
                // a side on this multibody fills (or adds to) the row's own J / M^-1 J out of the staged jacobians; a
                // side on ANOTHER multibody fills the second half of the row from that multibody's jacobians and LU
                const int mbA = l1 >= 0 ? mb_of_link[l1] : -1, mbB = l2 >= 0 ? mb_of_link[l2] : -1;
                int other = -1;
                float* J2 = WJ + R.nd_stride;
                float* WJ2 = J2 + R.nd_stride;
                auto link_side = [&](int link, int mbx, Vec3 center, Vec3 d) {
                    if (mbx == (int)m) {
                        mb_fill_geometry(V, M, (uint32_t)link - M.first_link, center, false, d, tmp, J, WJ, !first);
                        first = false;
                    } else {
                        const MbMeta MX = Vg.meta[mbx];
                        mb_fill_geometry(Vg, MX, (uint32_t)link - MX.first_link, center, false, d, tmp, J2, WJ2, false);
                        other = mbx;
                    }
                };
                if (l1 >= 0) link_side(l1, mbA, center1, dir);
                else kinematic_side(mf.body1, center1, dir);
                if (l2 >= 0) link_side(l2, mbB, center2, -dir);
                else kinematic_side(mf.body2, center2, -dir);
                if (first)  // cannot happen: the owner is one of the two
                    for (uint32_t cc = 0; cc < nd; ++cc) J[cc] = WJ[cc] = 0.f;
                float inv_r = mb_dot((int)nd, J, WJ);
                out_vel += mb_dot((int)nd, J, vel);
                out_vel += mb_dot((int)nd, J, ext);
                if (other >= 0) {
                    const MbMeta MX = Vg.meta[other];
                    inv_r += mb_dot((int)MX.ndofs, J2, WJ2);
                    out_vel += mb_dot((int)MX.ndofs, J2, Vg.vel + MX.dof_off);
                    out_vel += mb_dot((int)MX.ndofs, J2, Vg.ext + MX.dof_off);
                }
                MbRow row;
                row.r = inv_r != 0.f ? 1.f / inv_r : 1.f;
                row.other = other;
                row.two_sided = (l1 >= 0 && l2 >= 0) ? 1 : 0;
                row.contact = ci;
                row.slot = w;
                row.lim = mf.friction;
                row.dep = (int)(no0 + k - base);
                if (w == 0) {
                    float rhs = out_vel;
                    if (rhs <= -restitution_threshold) rhs += mf.restitution * rhs;
                    const float depth = c.depth + mf.margin1 + mf.margin2;
                    if (depth < 0.f) rhs += (-depth) * inv_dt;
                    row.rhs = rhs;
                    row.imp = cached.x * warmstart_coeff;
                    row.kind = NB2_ROW_UNILATERAL;
                    // SignoriniModel (signorini_model.rs:200-298): a row per ACTIVE contact only (is_constraint_active,
                    // :141-150); an inactive one keeps its cache entry as it is
                    if (model == NB2_CONTACT_SIGNORINI && !(depth >= 0.f)) {
                        row.kind = NB2_ROW_NONE;
                        row.imp = cached.x;
                    }
                    R.cpos[no0 + k] = xyz_f4(quat_inv_rotate(q1, normal), 0.f);
                } else {
                    row.rhs = out_vel;
                    row.imp = (w == 1 ? cached.y : cached.z) * warmstart_coeff;
                    row.kind = NB2_ROW_DEPENDENT;
                    if (model == NB2_CONTACT_SIGNORINI) {  // frictionless
                        row.kind = NB2_ROW_NONE;
                        row.imp = 0.f;
                    }
                }
                R.rows[rowi] = row;
            }
        }
    }
    if (lane != 0) return;  // the few unit-joint rows: one lane
    // ---- internal rows: unilateral (limits) first, then bilateral (motors), each in link order
    uint32_t iu = in0, ib = in0 + n_uni;
    for (uint32_t i = 0; i < M.n_links; ++i) {
        const MbLinkDev& l = L[i];
        if (l.type != NB2_MBJ_REVOLUTE && l.type != NB2_MBJ_PRISMATIC) continue;
        const float joint_velocity = vel[l.assembly];
        bool min_active = false;
        uint32_t min_row = 0;
        if (l.flags & NB2_MBJ_FLAG_MOTOR) {
            float* J = R.jw + (size_t)ib * rs;
            float* WJ = J + R.nd_stride;
            for (uint32_t c = 0; c < nd; ++c) J[c] = WJ[c] = 0.f;
            J[l.assembly] = 1.f;
            WJ[l.assembly] = 1.f;
            mb_lu_solve(V.mass + M.mass_off, V.piv + M.dof_off, (int)nd, WJ);
            MbRow row;
            const float v = l.motor_velocity > -l.motor_max_velocity ? (l.motor_velocity < l.motor_max_velocity ? l.motor_velocity : l.motor_max_velocity)
                                                                     : -l.motor_max_velocity;
            row.rhs = (joint_velocity + ext[l.assembly]) - v;
            row.r = 1.f / WJ[l.assembly];
            row.imp = l.impulses[0] * warmstart_coeff;
            row.lim = l.motor_max_force;
            row.kind = NB2_ROW_BILATERAL;
            row.dep = 0;
            row.other = -1;
            row.two_sided = 0;
            row.contact = 0xFFFFFFFFu;
            row.slot = (int)(i * 3 + 0);
            R.rows[ib++] = row;
        }
        if ((l.flags & NB2_MBJ_FLAG_MIN) && l.min_pos - l.coord >= 0.f) {
            float* J = R.jw + (size_t)iu * rs;
            float* WJ = J + R.nd_stride;
            for (uint32_t c = 0; c < nd; ++c) J[c] = WJ[c] = 0.f;
            J[l.assembly] = 1.f;
            WJ[l.assembly] = 1.f;
            mb_lu_solve(V.mass + M.mass_off, V.piv + M.dof_off, (int)nd, WJ);
            MbRow row;
            row.rhs = joint_velocity + ext[l.assembly];
            row.r = 1.f / WJ[l.assembly];
            row.imp = l.impulses[1] * warmstart_coeff;
            row.lim = 0.f;
            row.kind = NB2_ROW_UNILATERAL;
            row.dep = 0;
            row.other = -1;
            row.two_sided = 0;
            row.contact = 0xFFFFFFFFu;
            row.slot = (int)(i * 3 + 1);
            min_active = true;
            min_row = iu;
            R.rows[iu++] = row;
        }
        if ((l.flags & NB2_MBJ_FLAG_MAX) && -(l.max_pos - l.coord) >= 0.f) {
            float* J = R.jw + (size_t)iu * rs;
            float* WJ = J + R.nd_stride;
            for (uint32_t c = 0; c < nd; ++c) J[c] = WJ[c] = 0.f;
            J[l.assembly] = -1.f;
            if (min_active) {
                const float* W0 = R.jw + (size_t)min_row * rs + R.nd_stride;
                for (uint32_t c = 0; c < nd; ++c) WJ[c] = -W0[c];
            } else {
                WJ[l.assembly] = -1.f;
                mb_lu_solve(V.mass + M.mass_off, V.piv + M.dof_off, (int)nd, WJ);
            }
            MbRow row;
            row.rhs = -joint_velocity - ext[l.assembly];
            row.r = 1.f / (-WJ[l.assembly]);
            row.imp = l.impulses[2] * warmstart_coeff;
            row.lim = 0.f;
            row.kind = NB2_ROW_UNILATERAL;
            row.dep = 0;
            row.other = -1;
            row.two_sided = 0;
            row.contact = 0xFFFFFFFFu;
            row.slot = (int)(i * 3 + 2);
            R.rows[iu++] = row;
        }
    }
}

// rows a multibody will need: 3 per contact + its motors and limits (upper bound for the limits)
__global__ void k_mb_row_counts(MbView V, MbContacts C, uint32_t* row_cnt) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m == V.n_mb) row_cnt[m] = 0u;  // the scan runs over n_mb + 1 entries
    if (m >= V.n_mb) return;
    const MbMeta M = V.meta[m];
    const uint32_t nm = min(C.mcount[m], (uint32_t)NB2_MB_MANIFOLD_CAP);
    uint32_t n = 0;
    for (uint32_t a = 0; a < nm; ++a) n += 3 * C.manifolds[C.mlist[(size_t)m * NB2_MB_MANIFOLD_CAP + a]].num_contacts;
    const MbLinkDev* L = V.links + M.first_link;
    for (uint32_t i = 0; i < M.n_links; ++i)
        if (L[i].type == NB2_MBJ_REVOLUTE || L[i].type == NB2_MBJ_PRISMATIC)
            n += ((L[i].flags & NB2_MBJ_FLAG_MOTOR) ? 1 : 0) + ((L[i].flags & NB2_MBJ_FLAG_MIN) ? 1 : 0) + ((L[i].flags & NB2_MBJ_FLAG_MAX) ? 1 : 0);
    row_cnt[m] = n;
}

// SORProx::solve restricted to one multibody (sor_prox.rs:48-80, 159-179), one thread per multibody; then
// cache_impulses (signorini_coulomb_pyramid_model.rs:226-261; unit_joint rows: multibody.rs:1046-1053), the velocity
// update and Body::integrate (moreau_jean_solver.rs:328-347, multibody.rs:807-814).
#define NB2_MB_COMP_DOFS 1024   // generalized coordinates of one component (multibodies joined by manifolds)
#define NB2_MB_COMP_MEMBERS 48  // multibodies of one component
__global__ void __launch_bounds__(32 * MB_WPB) k_mb_velocity_solve(MbView V, MbRows R, MbCache K, MbContacts C,
                                                                   const nb2_contact* __restrict__ contacts, int iters, float dt) {
    // One warp per COMPONENT (the warp of its smallest member; a multibody that only touches the ground is its own
    // component).  mj_lambda of all members lives in the warp's slice of shared memory; a row is two coalesced loads
    // per side (J, M^-1 J), a butterfly reduction for J . mj_lambda and an axpy.  Sweep order = the reference's:
    // friction rows of the contacts between two links (manifold order), friction rows against the ground, every
    // member's unit-joint rows, then the normal rows in the same two classes (sor_prox.rs:159-179).  Only the summation
    // order inside a dot product differs from the one-thread form.
    __shared__ float s_lam[MB_WPB][NB2_MB_COMP_DOFS];
    __shared__ uint32_t s_mem[MB_WPB][NB2_MB_COMP_MEMBERS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= V.n_mb) return;
    if (R.comp[m] != m) return;  // solved by the warp of the component's smallest member
    float* lam = s_lam[warp];
    uint32_t* mem = s_mem[warp];
    const uint32_t nmem = R.coff[m + 1] - R.coff[m];
    const size_t rs = (size_t)4 * R.nd_stride;
    // ---- members in index order, their mj_lambda slots, the ordered contact list
    uint32_t ndtot = 0, ncont = 0, cstart = 0;
    if (lane == 0) {
        bool fits = nmem <= NB2_MB_COMP_MEMBERS;
        if (fits) {
            for (uint32_t k = 0; k < nmem; ++k) mem[k] = R.cmem[R.coff[m] + k];
            for (uint32_t a = 1; a < nmem; ++a) {
                const uint32_t v = mem[a];
                uint32_t b = a;
                while (b > 0 && mem[b - 1] > v) {
                    mem[b] = mem[b - 1];
                    --b;
                }
                mem[b] = v;
            }
            for (uint32_t k = 0; k < nmem; ++k) {
                R.lslot[mem[k]] = ndtot;
                ndtot += V.meta[mem[k]].ndofs;
                ncont += R.seg[4 * mem[k] + 3];
            }
            fits = ndtot <= NB2_MB_COMP_DOFS;
        }
        if (!fits) {
            atomicOr(&C.flags[0], NB2_FLAG_MB_OVERFLOW);
            ndtot = 0xFFFFFFFFu;
        } else {
            cstart = atomicAdd(R.cursor, ncont);
            // contacts by class (two links first, then against the ground), each class in manifold order: gather the
            // members' (sorted) manifold lists, sort by manifold index, expand to (normal row, first friction row)
            uint2* out = R.corder + cstart;
            uint2* tm = R.morder + cstart;  // (manifold index, member << 8 | position in the member's list); <= ncont entries
            uint32_t n = 0;
            for (int cls = 1; cls >= 0; --cls) {
                uint32_t nmf = 0;
                for (uint32_t k = 0; k < nmem; ++k) {
                    const uint32_t x = mem[k];
                    if (R.seg[4 * x + 3] == 0) continue;  // no contact rows (or they were dropped)
                    const uint32_t* list = C.mlist + (size_t)x * NB2_MB_MANIFOLD_CAP;
                    const uint32_t nm = min(C.mcount[x], (uint32_t)NB2_MB_MANIFOLD_CAP);
                    for (uint32_t a = 0; a < nm; ++a) {
                        const nb2_manifold& mf = C.manifolds[list[a]];
                        const int two = (C.link_of_body[mf.body1] >= 0 && C.link_of_body[mf.body2] >= 0) ? 1 : 0;
                        if (two != cls) continue;
                        uint32_t pos = nmf++;
                        while (pos > 0 && tm[pos - 1].x > list[a]) {
                            tm[pos] = tm[pos - 1];
                            --pos;
                        }
                        tm[pos] = make_uint2(list[a], (k << 8) | a);
                    }
                }
                for (uint32_t i = 0; i < nmf; ++i) {
                    const uint32_t x = mem[tm[i].y >> 8], a = tm[i].y & 255u;
                    const uint32_t* list = C.mlist + (size_t)x * NB2_MB_MANIFOLD_CAP;
                    const uint32_t* sg = R.seg + 4 * x;
                    const uint32_t basex = R.row_off[x], no0 = basex + sg[0] + sg[1] + sg[2];
                    uint32_t kk = 0;
                    for (uint32_t a2 = 0; a2 < a; ++a2) kk += C.manifolds[list[a2]].num_contacts;
                    const uint32_t nq = C.manifolds[list[a]].num_contacts;
                    for (uint32_t q = 0; q < nq && n < ncont; ++q) out[n++] = make_uint2(no0 + kk + q, basex + 2 * (kk + q));
                }
            }
            ncont = n;
        }
    }
    ndtot = __shfl_sync(0xFFFFFFFFu, ndtot, 0);
    if (ndtot == 0xFFFFFFFFu) return;
    ncont = __shfl_sync(0xFFFFFFFFu, ncont, 0);
    cstart = __shfl_sync(0xFFFFFFFFu, cstart, 0);
    __syncwarp();
    const uint2* order = R.corder + cstart;
    if (nmem == 1) {
        // ---- the common case, a multibody on its own: mj_lambda in registers (lane c holds [c] and [c + 32])
        const MbMeta M = V.meta[m];
        const int nd = (int)M.ndofs;
        const uint32_t* sg = R.seg + 4 * m;
        const uint32_t iu0 = R.row_off[m] + sg[0], iend = iu0 + sg[1] + sg[2];
        const bool h0 = (int)lane < nd, h1 = (int)lane + 32 < nd;
        float lam0 = 0.f, lam1 = 0.f;
        auto warm1 = [&](uint32_t r) {
            const float imp = R.rows[r].kind == NB2_ROW_NONE ? 0.f : R.rows[r].imp;
            if (imp != 0.f) {
                const float* WJ = R.jw + (size_t)r * rs + R.nd_stride;
                if (h0) lam0 = imp * WJ[lane] + lam0;
                if (h1) lam1 = imp * WJ[lane + 32] + lam1;
            }
        };
        auto solve1 = [&](uint32_t r) {
            const MbRow row = R.rows[r];
            if (row.kind == NB2_ROW_NONE) return;
            const float* J = R.jw + (size_t)r * rs;
            const float* WJ = J + R.nd_stride;
            const float j0 = h0 ? J[lane] : 0.f, j1 = h1 ? J[lane + 32] : 0.f;
            const float w0 = h0 ? WJ[lane] : 0.f, w1 = h1 ? WJ[lane + 32] : 0.f;
            float lo, hi;
            if (row.kind == NB2_ROW_UNILATERAL) {
                lo = 0.f;
                hi = NB2_F32_MAX;
            } else if (row.kind == NB2_ROW_BILATERAL) {
                lo = -row.lim;
                hi = row.lim;
            } else {  // Dependent (sor_prox.rs:251-272)
                const float dep = R.rows[R.row_off[m] + (uint32_t)row.dep].imp;
                if (dep == 0.f) {
                    if (row.imp != 0.f) {
                        lam0 = (-row.imp) * w0 + lam0;
                        lam1 = (-row.imp) * w1 + lam1;
                        if (lane == 0) R.rows[r].imp = 0.f;
                    }
                    return;
                }
                hi = row.lim * dep;
                lo = -hi;
            }
            float d = j0 * lam0 + j1 * lam1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
            d += row.rhs;
            float ni;
            if (row.kind == NB2_ROW_UNILATERAL) ni = fmaxf(0.f, row.imp - row.r * d);
            else {
                const float v = row.imp - row.r * d;
                ni = v > lo ? (v < hi ? v : hi) : lo;
            }
            const float dl = ni - row.imp;
            if (lane == 0) R.rows[r].imp = ni;
            lam0 = dl * w0 + lam0;
            lam1 = dl * w1 + lam1;
        };
        for (uint32_t q = 0; q < ncont; ++q) warm1(order[q].x);
        for (uint32_t q = 0; q < ncont; ++q) {
            warm1(order[q].y);
            warm1(order[q].y + 1);
        }
        for (uint32_t r = iu0; r < iend; ++r) warm1(r);
        for (int it = 0; it < iters; ++it) {
            for (uint32_t q = 0; q < ncont; ++q) {
                solve1(order[q].y);
                solve1(order[q].y + 1);
            }
            __syncwarp();
            for (uint32_t r = iu0; r < iend; ++r) solve1(r);
            for (uint32_t q = 0; q < ncont; ++q) solve1(order[q].x);
            __syncwarp();  // the normal impulses written by lane 0 are what the next sweep's friction rows read
        }
        if (h0) lam[lane] = lam0;
        if (h1) lam[lane + 32] = lam1;
        __syncwarp();
    } else {
    for (uint32_t c = lane; c < ndtot; c += 32) lam[c] = 0.f;
    __syncwarp();
    // ---- one row
    auto sides = [&](uint32_t r, const MbRow& row, uint32_t owner, float dl) {  // mj_lambda += dl * M^-1 J on both sides
        const uint32_t so = R.lslot[owner], no = V.meta[owner].ndofs;
        const float* WJ = R.jw + (size_t)r * rs + R.nd_stride;
        for (uint32_t c = lane; c < no; c += 32) lam[so + c] = dl * WJ[c] + lam[so + c];
        if (row.other >= 0) {
            const uint32_t s2 = R.lslot[row.other], n2 = V.meta[row.other].ndofs;
            const float* WJ2 = WJ + 2 * R.nd_stride;
            for (uint32_t c = lane; c < n2; c += 32) lam[s2 + c] = dl * WJ2[c] + lam[s2 + c];
        }
        __syncwarp();
    };
    auto warm = [&](uint32_t r, uint32_t owner) {
        const MbRow row = R.rows[r];
        if (row.kind != NB2_ROW_NONE && row.imp != 0.f) sides(r, row, owner, row.imp);
    };
    auto solve = [&](uint32_t r, uint32_t owner) {
        const MbRow row = R.rows[r];
        if (row.kind == NB2_ROW_NONE) return;
        float lo, hi;
        if (row.kind == NB2_ROW_UNILATERAL) {
            lo = 0.f;
            hi = NB2_F32_MAX;
        } else if (row.kind == NB2_ROW_BILATERAL) {
            lo = -row.lim;
            hi = row.lim;
        } else {  // Dependent (sor_prox.rs:251-272)
            const float dep = R.rows[R.row_off[owner] + (uint32_t)row.dep].imp;
            if (dep == 0.f) {
                if (row.imp != 0.f) {
                    sides(r, row, owner, -row.imp);
                    if (lane == 0) R.rows[r].imp = 0.f;
                }
                return;
            }
            hi = row.lim * dep;
            lo = -hi;
        }
        const uint32_t so = R.lslot[owner], no = V.meta[owner].ndofs;
        const float* J = R.jw + (size_t)r * rs;
        float d = 0.f;
        for (uint32_t c = lane; c < no; c += 32) d += J[c] * lam[so + c];
        if (row.other >= 0) {
            const uint32_t s2 = R.lslot[row.other], n2 = V.meta[row.other].ndofs;
            const float* J2 = J + 2 * R.nd_stride;
            for (uint32_t c = lane; c < n2; c += 32) d += J2[c] * lam[s2 + c];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
        d += row.rhs;
        float ni;
        if (row.kind == NB2_ROW_UNILATERAL) ni = fmaxf(0.f, row.imp - row.r * d);
        else {
            const float v = row.imp - row.r * d;
            ni = v > lo ? (v < hi ? v : hi) : lo;
        }
        if (lane == 0) R.rows[r].imp = ni;
        sides(r, row, owner, ni - row.imp);
    };
    auto owner_of = [&](uint32_t r) -> uint32_t {  // the member whose row range holds r
        uint32_t x = mem[0];
        for (uint32_t k = 1; k < nmem; ++k)
            if (R.row_off[mem[k]] <= r) x = mem[k];
        return x;
    };
    auto internal = [&](bool warm_start) {
        for (uint32_t k = 0; k < nmem; ++k) {
            const uint32_t x = mem[k];
            const uint32_t* sg = R.seg + 4 * x;
            const uint32_t iu0 = R.row_off[x] + sg[0];
            for (uint32_t r = iu0; r < iu0 + sg[1] + sg[2]; ++r) {
                if (warm_start) warm(r, x);
                else solve(r, x);
            }
        }
    };
    // warmstart_set: unilateral rows, then bilateral rows, then the internal ones (sor_prox.rs:57-65)
    for (uint32_t q = 0; q < ncont; ++q) warm(order[q].x, nmem == 1 ? m : owner_of(order[q].x));
    for (uint32_t q = 0; q < ncont; ++q) {
        const uint32_t x = nmem == 1 ? m : owner_of(order[q].x);
        warm(order[q].y, x);
        warm(order[q].y + 1, x);
    }
    internal(true);
    for (int it = 0; it < iters; ++it) {
        for (uint32_t q = 0; q < ncont; ++q) {  // step_bilateral(contacts)
            const uint32_t x = nmem == 1 ? m : owner_of(order[q].x);
            solve(order[q].y, x);
            solve(order[q].y + 1, x);
        }
        internal(false);                         // unilateral_ground, bilateral_ground of every member
        for (uint32_t q = 0; q < ncont; ++q) solve(order[q].x, nmem == 1 ? m : owner_of(order[q].x));  // step_unilateral(contacts)
    }
    }
    __syncwarp();
    // ---- cache_impulses, velocity update, Body::integrate of every member
    for (uint32_t q = lane; q < ncont; q += 32) {
        const uint32_t rn = order[q].x, rf = order[q].y;
        const uint32_t ci = R.rows[rn].contact;
        K.imp_cur[ci] = make_float4(R.rows[rn].imp, R.rows[rf].imp, R.rows[rf + 1].imp, 0.f);
        K.ckey_cur[ci] = contacts[ci].key;
    }
    for (uint32_t k = 0; k < nmem; ++k) {
        const uint32_t x = mem[k];
        const MbMeta MX = V.meta[x];
        MbLinkDev* L = V.links + MX.first_link;
        const uint32_t* sg = R.seg + 4 * x;
        const uint32_t iu0 = R.row_off[x] + sg[0];
        if (lane == 0)
            for (uint32_t r = iu0; r < iu0 + sg[1] + sg[2]; ++r) {
                const int sl = R.rows[r].slot;
                L[sl / 3].impulses[sl % 3] = R.rows[r].imp;
            }
        float* vel = V.vel + MX.dof_off;
        const float* ext = V.ext + MX.dof_off;
        float* glam = V.lam + MX.dof_off;
        const uint32_t so = R.lslot[x];
        for (uint32_t c = lane; c < MX.ndofs; c += 32) {
            const float l = lam[so + c];
            glam[c] = l;
            float v = vel[c];
            v += ext[c];
            v += l;
            vel[c] = v;
        }
        __syncwarp();
        for (uint32_t i = lane; i < MX.n_links; i += 32) mbj_integrate(L[i], dt, vel + L[i].assembly);
    }
}

// Body::apply_displacement (multibody.rs:816-824): the joints take their share, the kinematics are redone.  A real
// call (not inlined): the caller re-reads the poses it wrote from memory.
__device__ __noinline__ void mb_displace(const MbView& V, const MbMeta& MX, const Proxies& P, const float* disp) {
    MbLinkDev* L = V.links + MX.first_link;
    for (uint32_t i = 0; i < MX.n_links; ++i) mbj_apply_displacement(L[i], disp + L[i].assembly);
    mb_update_kinematics<1>(V, MX, P, 0u);
}

// NonlinearSORProx::solve restricted to one multibody (nonlinear_sor_prox.rs:17-55): per iteration its internal
// position constraints (multibody.rs:1112-1152, unit_joint.rs:198-251), then its contacts in manifold order
// (update_contact_constraint :156-294, solve_unilateral :121-154); every displacement re-runs update_kinematics.
__global__ void __launch_bounds__(MB_TPB) k_mb_position_solve(MbView V, Proxies P, MbContacts C, MbRows R, PosParams PP, int iters) {
    // one thread per component (the thread of its smallest member), see k_mb_velocity_solve
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= V.n_mb) return;
    if (R.comp[m] != m) return;
    const uint32_t nmem = min(R.coff[m + 1] - R.coff[m], (uint32_t)NB2_MB_COMP_MEMBERS);
    uint32_t mem[NB2_MB_COMP_MEMBERS];
    unsigned char ptr[NB2_MB_COMP_MEMBERS];
    unsigned short kc[NB2_MB_COMP_MEMBERS];
    for (uint32_t k = 0; k < nmem; ++k) mem[k] = R.cmem[R.coff[m] + k];
    for (uint32_t a = 1; a < nmem; ++a) {
        const uint32_t v = mem[a];
        uint32_t b = a;
        while (b > 0 && mem[b - 1] > v) {
            mem[b] = mem[b - 1];
            --b;
        }
        mem[b] = v;
    }
    float J[NB2_MB_MAX_DOFS], WJ[NB2_MB_MAX_DOFS], tmp[NB2_MB_MAX_DOFS];
    auto displace = [&](const MbMeta& MX, const float* disp) { mb_displace(V, MX, P, disp); };
    for (int it = 0; it < iters; ++it) {
        // step_solve_internal_position_constraints of every member with unit-joint constraints (nonlinear_sor_prox.rs:40-44)
        for (uint32_t k = 0; k < nmem; ++k) {
            const MbMeta M = V.meta[mem[k]];
            if (!M.has_internal) continue;
            const uint32_t nd = M.ndofs;
            MbLinkDev* L = V.links + M.first_link;
            for (uint32_t i = 0; i < M.n_links; ++i) {
                const MbLinkDev& l = L[i];
                if (l.type != NB2_MBJ_REVOLUTE && l.type != NB2_MBJ_PRISMATIC) continue;
                if (!(l.flags & (NB2_MBJ_FLAG_MIN | NB2_MBJ_FLAG_MAX))) continue;
                float sign = 1.f, rhs = 0.f;
                bool some = false;
                if (l.flags & NB2_MBJ_FLAG_MIN) {
                    const float err = l.min_pos - l.coord;
                    if (err > 0.f) {
                        rhs = -err;
                        some = true;
                    }
                }
                if (!some && (l.flags & NB2_MBJ_FLAG_MAX)) {
                    const float err = -(l.max_pos - l.coord);
                    if (err > 0.f) {
                        rhs = -err;
                        some = true;
                        sign = -1.f;
                    }
                }
                if (!some) continue;
                for (uint32_t c = 0; c < nd; ++c) WJ[c] = 0.f;
                WJ[l.assembly] = sign;
                mb_lu_solve(V.mass + M.mass_off, V.piv + M.dof_off, (int)nd, WJ);
                const float r = 1.f / (sign * WJ[l.assembly]);
                const float crhs = clamp_rhs(rhs, l.type == NB2_MBJ_REVOLUTE, PP);
                if (crhs < 0.f) {
                    const float impulse = -crhs * r;
                    for (uint32_t c = 0; c < nd; ++c) WJ[c] *= impulse;
                    displace(M, WJ);
                }
            }
            mb_update_kinematics<1>(V, M, P, 0u);
        }
        // the component's contacts in manifold order: a merge of the members' (sorted) manifold lists
        for (uint32_t k = 0; k < nmem; ++k) {
            ptr[k] = 0;
            kc[k] = 0;
        }
        for (;;) {
            uint32_t best = 0xFFFFFFFFu, bk = 0;
            for (uint32_t k = 0; k < nmem; ++k) {
                const uint32_t x = mem[k];
                if (R.seg[4 * x + 3] == 0) continue;
                const uint32_t nm = min(C.mcount[x], (uint32_t)NB2_MB_MANIFOLD_CAP);
                if (ptr[k] >= nm) continue;
                const uint32_t mi = C.mlist[(size_t)x * NB2_MB_MANIFOLD_CAP + ptr[k]];
                if (mi < best) {
                    best = mi;
                    bk = k;
                }
            }
            if (best == 0xFFFFFFFFu) break;
            ++ptr[bk];
            const uint32_t x = mem[bk];
            const uint32_t* sg = R.seg + 4 * x;
            const uint32_t no0 = R.row_off[x] + sg[0] + sg[1] + sg[2];
            const nb2_manifold& mf = C.manifolds[best];
            const int l1 = C.link_of_body[mf.body1], l2 = C.link_of_body[mf.body2];
            const Pose c1 = load_coll(mf.coll1_wrt_body), c2 = load_coll(mf.coll2_wrt_body);
            for (uint32_t ci = mf.first_contact; ci < mf.first_contact + mf.num_contacts && ci < C.n_contacts; ++ci) {
                const uint32_t k = kc[bk]++;
                if (R.rows[no0 + k].kind == NB2_ROW_NONE) continue;  // SignoriniModel: no position constraint for an inactive contact (signorini_model.rs:230-232)
                const nb2_contact& c = C.contacts[ci];
                Pose b1, b2;
                b1.t = f4_xyz(P.pos_t[mf.body1]);
                b1.r = f4_quat(P.pos_q[mf.body1]);
                b2.t = f4_xyz(P.pos_t[mf.body2]);
                b2.r = f4_quat(P.pos_q[mf.body2]);
                const Pose m1 = pose_mul(b1, c1), m2 = pose_mul(b2, c2);
                const float4 n1 = R.cpos[no0 + k];
                ContactEval cev;
                if (!kinematic_contact(make_float4(c.local1[0], c.local1[1], c.local1[2], c.dilation1 + mf.margin1),
                                       make_float4(c.local2[0], c.local2[1], c.local2[2], c.dilation2 + mf.margin2),
                                       make_float4(c.dir1[0], c.dir1[1], c.dir1[2], __int_as_float((int)c.geom1)),
                                       make_float4(c.dir2[0], c.dir2[1], c.dir2[2], __int_as_float((int)c.geom2)), n1, m1, m2, &cev))
                    continue;
                const float rhs = clamp_rhs(-cev.depth, false, PP);
                if (rhs >= 0.f) continue;
                // the reference displaces body 1 and then body 2 with jacobians taken BEFORE either moved
                // (nonlinear_sor_prox.rs:131-152); they may be two links of one multibody or of two
                float inv_r = 0.f;
                float J2[NB2_MB_MAX_DOFS], WJ2[NB2_MB_MAX_DOFS];
                const int xa = l1 >= 0 ? C.mb_of_link[l1] : -1, xb = l2 >= 0 ? C.mb_of_link[l2] : -1;
                if (xa >= 0) {
                    const MbMeta M1 = V.meta[xa];
                    mb_fill_geometry(V, M1, (uint32_t)l1 - M1.first_link, cev.world1, false, -cev.normal, tmp, J, WJ, false);
                    inv_r += mb_dot((int)M1.ndofs, J, WJ);
                }
                if (xb >= 0) {
                    const MbMeta M2 = V.meta[xb];
                    mb_fill_geometry(V, M2, (uint32_t)l2 - M2.first_link, cev.world2, false, cev.normal, tmp, J2, WJ2, false);
                    inv_r += mb_dot((int)M2.ndofs, J2, WJ2);
                }
                if (inv_r == 0.f) continue;
                const float impulse = -rhs * (1.f / inv_r);
                if (xa >= 0) {
                    const MbMeta M1 = V.meta[xa];
                    for (uint32_t cc = 0; cc < M1.ndofs; ++cc) WJ[cc] *= impulse;
                    displace(M1, WJ);
                }
                if (xb >= 0) {
                    const MbMeta M2 = V.meta[xb];
                    for (uint32_t cc = 0; cc < M2.ndofs; ++cc) WJ2[cc] *= impulse;
                    displace(M2, WJ2);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
__global__ void k_mb_init_links(MbLinkDev* links, uint32_t n_links, const nb2_body* __restrict__ raw, uint32_t n_bodies, int* link_of_body,
                                uint32_t* flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_links) return;
    MbLinkDev& l = links[i];
    if (l.body < 0 || (uint32_t)l.body >= n_bodies || raw[l.body].status != NB2_BODY_MULTIBODY_LINK) {
        atomicOr(&flags[0], NB2_FLAG_MB_COUPLED << 2);  // bad link record
        l.body = -1;
        return;
    }
    const nb2_body& b = raw[l.body];
    for (int k = 0; k < 3; ++k) l.local_com[k] = b.local_com[k];
    l.mass = b.mass;
    for (int k = 0; k < 9; ++k) l.local_inertia[k] = b.local_inertia[k];
    link_of_body[l.body] = (int)i;
}
__global__ void k_mb_pack_links(const MbLinkDev* __restrict__ links, uint32_t n_links, const float* __restrict__ vel,
                                const MbMeta* __restrict__ meta, const int* __restrict__ mb_of_link, nb2_mb_link* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_links) return;
    const MbLinkDev& l = links[i];
    nb2_mb_link& o = out[i];
    switch (l.type) {
        case NB2_MBJ_FREE:
        case NB2_MBJ_FIXED:
            for (int k = 0; k < 3; ++k) o.coords[k] = l.free_t[k];
            for (int k = 0; k < 4; ++k) o.coords[3 + k] = l.free_q[k];
            break;
        case NB2_MBJ_BALL:
            for (int k = 0; k < 4; ++k) o.coords[k] = l.rot[k];
            break;
        default: o.coords[0] = l.coord; break;
    }
    const float* v = vel + meta[mb_of_link[i]].dof_off + l.assembly;
    for (uint32_t k = 0; k < l.ndofs; ++k) o.velocity[k] = v[k];
    if (l.ndofs == 1)
        for (int k = 0; k < 3; ++k) o.impulses[k] = l.impulses[k];
}

struct MbState {
    uint32_t n_mb = 0, n_links = 0, total_dofs = 0, nd_max = 0;
    DevBuf<MbMeta> meta;
    DevBuf<MbLinkDev> links;
    DevBuf<nb2_mb_link> recs;  // as uploaded (download template)
    DevBuf<int> mb_of_link, link_of_body, piv;
    DevBuf<float> vel, damp, acc, ext, lam, jac, cor, icd, mass, accw;
    DevBuf<uint32_t> mcount, mlist, row_cnt, row_off, seg;
    DevBuf<uint32_t> edges, ecount, comp, ccount, coff, ccursor, cmem, lslot;  // components; ecount: [0] edges, [1] the corder allocator
    DevBuf<uint2> corder, morder;
    DevBuf<MbRow> rows;
    DevBuf<float> jw;
    DevBuf<float4> cpos;
    uint32_t row_cap = 0;
    uint32_t link_bodies = 0;  // n_bodies the link_of_body map was sized for
    uint32_t jac_words = 0, stage_words = 0;  // largest body-jacobian block; per-thread staging region of k_mb_refresh (0 = none)
    int refresh_tpb = MB_TPB;
    bool assemble_attr = false;
};

static MbState* mb_state(Context* ctx) { return reinterpret_cast<MbState*>(ctx->mb); }

static MbView mb_view(MbState* S) {
    MbView V;
    V.meta = S->meta.p;
    V.links = S->links.p;
    V.vel = S->vel.p;
    V.damp = S->damp.p;
    V.acc = S->acc.p;
    V.ext = S->ext.p;
    V.lam = S->lam.p;
    V.jac = S->jac.p;
    V.cor = S->cor.p;
    V.icd = S->icd.p;
    V.mass = S->mass.p;
    V.accw = S->accw.p;
    V.piv = S->piv.p;
    V.n_mb = S->n_mb;
    return V;
}
static Proxies mb_proxies(Context* ctx) {
    Proxies P;
    P.pos_t = ctx->pos_t.p;
    P.pos_q = ctx->pos_q.p;
    P.com_im = ctx->com_im.p;
    P.vel = ctx->vel.p;
    return P;
}
static inline unsigned int mb_blocks(uint32_t n) { return (n + MB_TPB - 1) / MB_TPB; }

void mb_release(Context* ctx) {
    MbState* S = mb_state(ctx);
    if (!S) return;
    S->meta.release(); S->links.release(); S->recs.release(); S->mb_of_link.release(); S->link_of_body.release(); S->piv.release();
    S->vel.release(); S->damp.release(); S->acc.release(); S->ext.release(); S->lam.release(); S->jac.release(); S->cor.release();
    S->icd.release(); S->mass.release(); S->accw.release(); S->mcount.release(); S->mlist.release(); S->row_cnt.release(); S->row_off.release();
    S->seg.release(); S->rows.release(); S->jw.release(); S->cpos.release();
    S->edges.release(); S->ecount.release(); S->comp.release(); S->ccount.release(); S->coff.release(); S->ccursor.release();
    S->cmem.release(); S->lslot.release(); S->corder.release(); S->morder.release();
    delete S;
    ctx->mb = nullptr;
}

// a new body set drops the multibodies (their links name body records), like the joints
void mb_invalidate(Context* ctx) {
    if (mb_state(ctx)) mb_state(ctx)->n_mb = 0;
}

int mb_count(Context* ctx) { return mb_state(ctx) ? (int)mb_state(ctx)->n_mb : 0; }

struct MbState;
static int mb_refresh(Context* ctx, MbState* S, int mode);

int mb_upload(Context* ctx, const nb2_multibody* mbs, uint32_t n_mb, const nb2_mb_link* links, uint32_t n_links) {
    if (ctx->n_bodies == 0) return set_error(ctx, NB2_ERR_NOT_READY, "upload the bodies before the multibodies");
    if (!mb_state(ctx)) ctx->mb = new MbState();
    MbState* S = mb_state(ctx);
    S->n_mb = 0;
    if (n_mb == 0 || n_links == 0) return NB2_OK;
    if (!mbs || !links) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null multibody arrays");
    std::vector<MbMeta> meta(n_mb);
    std::vector<MbLinkDev> dev(n_links);
    std::vector<int> mb_of_link(n_links, -1);
    std::vector<float> vel, damp;
    uint32_t dof_off = 0, jac_off = 0, mass_off = 0, nd_max = 0, jac_max = 0;
    for (uint32_t m = 0; m < n_mb; ++m) {
        const nb2_multibody& r = mbs[m];
        if (r.n_links == 0 || (size_t)r.first_link + r.n_links > n_links) return set_error(ctx, NB2_ERR_BAD_INDEX, "multibody %u: link range", m);
        MbMeta M;
        memset(&M, 0, sizeof(M));
        M.first_link = r.first_link;
        M.n_links = r.n_links;
        M.flags = r.flags;
        M.dof_off = dof_off;
        M.jac_off = jac_off;
        M.mass_off = mass_off;
        uint32_t nd = 0;
        for (uint32_t k = 0; k < r.n_links; ++k) {
            const nb2_mb_link& s = links[r.first_link + k];
            if (s.joint_type >= NB2_MBJ_TYPE_COUNT) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "link %u: joint type", r.first_link + k);
            if (s.parent >= (int)k || (k == 0) != (s.parent < 0)) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "link %u: parent", r.first_link + k);
            if (s.multibody != (int)m || mb_of_link[r.first_link + k] != -1) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "link %u: multibody index", r.first_link + k);
            if (s.body < 0 || (uint32_t)s.body >= ctx->n_bodies) return set_error(ctx, NB2_ERR_BAD_INDEX, "link %u: body", r.first_link + k);
            mb_of_link[r.first_link + k] = (int)m;
            MbLinkDev& d = dev[r.first_link + k];
            memset(&d, 0, sizeof(d));
            d.parent = s.parent;
            d.type = (int)s.joint_type;
            d.flags = s.flags;
            d.body = s.body;
            d.assembly = nd;
            static const uint32_t NDOFS[NB2_MBJ_TYPE_COUNT] = {6, 3, 1, 1, 0};
            d.ndofs = NDOFS[s.joint_type];
            for (int a = 0; a < 3; ++a) {
                d.parent_shift[a] = s.parent_shift[a];
                d.body_shift[a] = s.body_shift[a];
                d.axis[a] = s.axis[a];
                d.impulses[a] = d.ndofs == 1 ? s.impulses[a] : 0.f;
            }
            d.min_pos = s.min_pos;
            d.max_pos = s.max_pos;
            d.motor_velocity = s.motor_velocity;
            d.motor_max_velocity = s.motor_max_velocity;
            d.motor_max_force = s.motor_max_force;
            d.free_q[3] = 1.f;
            d.rot[3] = 1.f;
            switch (s.joint_type) {
                case NB2_MBJ_FREE:
                case NB2_MBJ_FIXED:
                    for (int a = 0; a < 3; ++a) d.free_t[a] = s.coords[a];
                    for (int a = 0; a < 4; ++a) d.free_q[a] = s.coords[3 + a];
                    break;
                case NB2_MBJ_BALL:
                    for (int a = 0; a < 4; ++a) d.rot[a] = s.coords[a];
                    break;
                case NB2_MBJ_REVOLUTE: {  // Rotation::from_axis_angle(axis, angle)
                    d.coord = s.coords[0];
                    const float h = d.coord * 0.5f, sn = sinf(h);
                    d.rot[0] = d.axis[0] * sn;
                    d.rot[1] = d.axis[1] * sn;
                    d.rot[2] = d.axis[2] * sn;
                    d.rot[3] = cosf(h);
                    break;
                }
                default: d.coord = s.coords[0]; break;
            }
            for (uint32_t a = 0; a < d.ndofs; ++a) {
                vel.push_back(s.velocity[a]);
                damp.push_back(s.damping[a]);
            }
            if (d.ndofs == 1 && (d.flags & (NB2_MBJ_FLAG_MIN | NB2_MBJ_FLAG_MAX | NB2_MBJ_FLAG_MOTOR))) M.has_internal = 1;
            nd += d.ndofs;
        }
        if (nd == 0 || nd > NB2_MB_MAX_DOFS) return set_error(ctx, NB2_ERR_UNSUPPORTED, "multibody %u: %u dofs (1..%d supported)", m, nd, NB2_MB_MAX_DOFS);
        M.ndofs = nd;
        nd_max = nd > nd_max ? nd : nd_max;
        jac_max = r.n_links * 6 * nd > jac_max ? r.n_links * 6 * nd : jac_max;
        dof_off += nd;
        jac_off += r.n_links * 6 * nd;
        mass_off += nd * nd;
        meta[m] = M;
    }
    for (uint32_t i = 0; i < n_links; ++i)
        if (mb_of_link[i] < 0) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "link %u belongs to no multibody", i);
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    S->n_links = n_links;
    S->total_dofs = dof_off;
    S->nd_max = nd_max;
    NB2_TRY(S->meta.reserve(ctx, n_mb));
    NB2_TRY(S->links.reserve(ctx, n_links));
    NB2_TRY(S->recs.reserve(ctx, n_links));
    NB2_TRY(S->mb_of_link.reserve(ctx, n_links));
    NB2_TRY(S->link_of_body.reserve(ctx, ctx->n_bodies));
    NB2_TRY(S->piv.reserve(ctx, dof_off));
    NB2_TRY(S->vel.reserve(ctx, dof_off));
    NB2_TRY(S->damp.reserve(ctx, dof_off));
    NB2_TRY(S->acc.reserve(ctx, dof_off));
    NB2_TRY(S->ext.reserve(ctx, dof_off));
    NB2_TRY(S->lam.reserve(ctx, dof_off));
    NB2_TRY(S->jac.reserve(ctx, jac_off));
    NB2_TRY(S->cor.reserve(ctx, jac_off));
    NB2_TRY(S->icd.reserve(ctx, (size_t)6 * dof_off));
    NB2_TRY(S->mass.reserve(ctx, mass_off));
    NB2_TRY(S->accw.reserve(ctx, (size_t)6 * n_links));
    NB2_TRY(S->mcount.reserve(ctx, n_mb));
    NB2_TRY(S->mlist.reserve(ctx, (size_t)n_mb * NB2_MB_MANIFOLD_CAP));
    NB2_TRY(S->row_cnt.reserve(ctx, n_mb + 1));
    NB2_TRY(S->row_off.reserve(ctx, n_mb + 2));
    NB2_TRY(S->seg.reserve(ctx, (size_t)4 * n_mb));
    NB2_TRY(S->ecount.reserve(ctx, 4));
    NB2_TRY(S->comp.reserve(ctx, n_mb));
    NB2_TRY(S->ccount.reserve(ctx, n_mb + 1));
    NB2_TRY(S->coff.reserve(ctx, n_mb + 2));
    NB2_TRY(S->ccursor.reserve(ctx, n_mb));
    NB2_TRY(S->cmem.reserve(ctx, n_mb));
    NB2_TRY(S->lslot.reserve(ctx, n_mb));
    S->link_bodies = ctx->n_bodies;
    NB2_CUDA(ctx, cudaMemcpyAsync(S->meta.p, meta.data(), n_mb * sizeof(MbMeta), cudaMemcpyHostToDevice, ctx->stream));
    NB2_CUDA(ctx, cudaMemcpyAsync(S->links.p, dev.data(), n_links * sizeof(MbLinkDev), cudaMemcpyHostToDevice, ctx->stream));
    NB2_CUDA(ctx, cudaMemcpyAsync(S->recs.p, links, n_links * sizeof(nb2_mb_link), cudaMemcpyHostToDevice, ctx->stream));
    NB2_CUDA(ctx, cudaMemcpyAsync(S->mb_of_link.p, mb_of_link.data(), n_links * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NB2_CUDA(ctx, cudaMemcpyAsync(S->vel.p, vel.data(), dof_off * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    NB2_CUDA(ctx, cudaMemcpyAsync(S->damp.p, damp.data(), dof_off * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(S->link_of_body.p, 0xFF, ctx->n_bodies * sizeof(int), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(S->seg.p, 0, (size_t)4 * n_mb * sizeof(uint32_t), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(S->mcount.p, 0, n_mb * sizeof(uint32_t), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(S->row_off.p, 0, (n_mb + 2) * sizeof(uint32_t), ctx->stream));
    NB2_TRY(ctx->flags.reserve(ctx, 4));
    k_mb_init_links<<<mb_blocks(n_links), MB_TPB, 0, ctx->stream>>>(S->links.p, n_links, ctx->raw.p, ctx->n_bodies, S->link_of_body.p, ctx->flags.p);
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors go out of scope
    S->n_mb = n_mb;
    {  // staging geometry of k_mb_refresh: as many threads per block as regions fit the opt-in shared memory
        S->jac_words = jac_max;
        uint32_t words = 2 * jac_max + nd_max * nd_max + 6 * nd_max + nd_max;
        words |= 1u;
        size_t optin = ctx->smem_optin ? ctx->smem_optin : 48 * 1024;
        int fit = (int)(optin / ((size_t)words * 4));
        if (fit >= 1) {
            S->stage_words = words;
            S->refresh_tpb = 32 * (fit > MB_WPB ? MB_WPB : fit);
            NB2_CUDA(ctx, cudaFuncSetAttribute(k_mb_refresh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin));
        } else {
            S->stage_words = 0;
            S->refresh_tpb = 32 * MB_WPB;
        }
    }
    // poses and velocities of the links' body records are valid from now on
    NB2_TRY(mb_refresh(ctx, S, 0));
    ctx->launches += 1;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

static int mb_refresh(Context* ctx, MbState* S, int mode) {
    const Vec3 g = mk3(ctx->params.gravity[0], ctx->params.gravity[1], ctx->params.gravity[2]);
    const bool stage = S->stage_words != 0 && mode != 2;
    const int tpb = stage ? S->refresh_tpb : 32 * MB_WPB;  // one warp per multibody
    const int wpb = tpb / 32;
    const size_t smem = stage ? (size_t)wpb * S->stage_words * 4 : 0;
    k_mb_refresh<<<(S->n_mb + wpb - 1) / wpb, tpb, smem, ctx->stream>>>(mb_view(S), mb_proxies(ctx), ctx->params.dt, g, mode,
                                                                      stage ? S->stage_words : 0u, S->jac_words);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int mb_download_links(Context* ctx, nb2_mb_link* out, uint32_t n) {
    MbState* S = mb_state(ctx);
    if (!S || n > S->n_links) return set_error(ctx, NB2_ERR_BAD_INDEX, "more links than uploaded");
    if (n == 0) return NB2_OK;
    k_mb_pack_links<<<mb_blocks(n), MB_TPB, 0, ctx->stream>>>(S->links.p, n, S->vel.p, S->meta.p, S->mb_of_link.p, S->recs.p);
    ctx->launches++;
    NB2_CUDA(ctx, cudaMemcpyAsync(out, S->recs.p, n * sizeof(nb2_mb_link), cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB2_OK;
}

// step stage 1 (mechanical_world.rs:230-243): kinematics, dynamics, accelerations; the links' body records follow
int mb_launch_refresh(Context* ctx) {
    MbState* S = mb_state(ctx);
    if (!S || S->n_mb == 0) return NB2_OK;
    return mb_refresh(ctx, S, 1);
}

static MbContacts mb_contacts(Context* ctx, MbState* S) {
    MbContacts C;
    C.manifolds = ctx->manifolds.p;
    C.contacts = ctx->contacts.p;
    C.n_manifolds = ctx->n_manifolds;
    C.n_contacts = ctx->n_contacts;
    C.link_of_body = S->link_of_body.p;
    C.mb_of_link = S->mb_of_link.p;
    C.status = ctx->b_status.p;
    C.pos_t = ctx->pos_t.p;
    C.pos_q = ctx->pos_q.p;
    C.vel = ctx->vel.p;
    C.com_im = ctx->com_im.p;
    C.mcount = S->mcount.p;
    C.mlist = S->mlist.p;
    C.edges = S->edges.p;
    C.n_edges = S->ecount.p;
    C.flags = ctx->flags.p;
    return C;
}
static MbRows mb_rows(MbState* S) {
    MbRows R;
    R.rows = S->rows.p;
    R.jw = S->jw.p;
    R.cpos = S->cpos.p;
    R.row_off = S->row_off.p;
    R.row_cnt = S->row_cnt.p;
    R.seg = S->seg.p;
    R.nd_stride = S->nd_max;
    R.row_cap = S->row_cap;
    R.comp = S->comp.p;
    R.coff = S->coff.p;
    R.cmem_c = S->cmem.p;
    R.cmem = S->cmem.p;
    R.lslot = S->lslot.p;
    R.corder = S->corder.p;
    R.morder = S->morder.p;
    R.cursor = S->ecount.p + 1;
    return R;
}

// step stages 2-4: rows, velocity resolution, impulse caching, velocity update and integration.  Runs AFTER the rigid
// path's launch_cache_impulses (which carries the multibody contacts' cache entries over as if they slept) so that
// this step's impulses are what the next step finds.
int mb_launch_velocity(Context* ctx) {
    MbState* S = mb_state(ctx);
    if (!S || S->n_mb == 0) return NB2_OK;
    const uint32_t n_mb = S->n_mb;
    // rows: 3 per contact of the world at most + 3 per unit-joint link
    const size_t cap = (size_t)3 * ctx->n_contacts + (size_t)3 * S->n_links + 4;
    if (cap > 0x7FFFFFFFull) return set_error(ctx, NB2_ERR_UNSUPPORTED, "too many multibody rows");
    NB2_TRY(S->rows.reserve(ctx, cap));
    NB2_TRY(S->jw.reserve(ctx, cap * 4 * S->nd_max));
    NB2_TRY(S->cpos.reserve(ctx, cap));
    NB2_TRY(S->corder.reserve(ctx, cap / 3 + 4));
    NB2_TRY(S->morder.reserve(ctx, cap / 3 + 4));
    NB2_TRY(S->edges.reserve(ctx, (size_t)2 * ctx->n_manifolds + 2));
    S->row_cap = (uint32_t)cap;
    MbView V = mb_view(S);
    MbContacts C = mb_contacts(ctx, S);
    NB2_CUDA(ctx, cudaMemsetAsync(S->mcount.p, 0, n_mb * sizeof(uint32_t), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(S->ecount.p, 0, 4 * sizeof(uint32_t), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(S->ccursor.p, 0, n_mb * sizeof(uint32_t), ctx->stream));
    if (ctx->n_manifolds)
        k_mb_collect<<<(ctx->n_manifolds + 127) / 128, 128, 0, ctx->stream>>>(C, S->mb_of_link.p);
    // components of the multibodies that share manifolds, their member lists
    k_mb_components<<<1, 256, 0, ctx->stream>>>(n_mb, S->edges.p, S->ecount.p, S->comp.p, S->ccount.p);
    NB2_TRY(exclusive_scan_u32(ctx, S->ccount.p, S->coff.p, n_mb + 1));
    k_mb_members<<<mb_blocks(n_mb), MB_TPB, 0, ctx->stream>>>(n_mb, S->comp.p, S->coff.p, S->ccursor.p, S->cmem.p);
    k_mb_row_counts<<<mb_blocks(n_mb + 1), MB_TPB, 0, ctx->stream>>>(V, C, S->row_cnt.p);
    NB2_TRY(exclusive_scan_u32(ctx, S->row_cnt.p, S->row_off.p, n_mb + 1));
    ctx->launches += 2;
    MbRows R = mb_rows(S);
    MbCache K;
    const int cur = ctx->cur, prev = 1 - cur;
    K.ckey_prev = ctx->ckey[prev].p;
    K.imp_prev = ctx->imp[prev].p;
    K.n_prev = ctx->imp_n[prev];
    K.imp_cur = ctx->imp[cur].p;
    K.ckey_cur = ctx->ckey[cur].p;
    {
        const uint32_t words = (S->jac_words + S->nd_max * S->nd_max + S->nd_max) | 1u;
        const size_t optin = ctx->smem_optin ? ctx->smem_optin : 48 * 1024;
        int fit = (int)(optin / ((size_t)words * 4));
        const bool stage = fit >= 1;
        const int wpb = stage ? (fit > MB_WPB ? MB_WPB : fit) : MB_WPB;
        if (stage && !S->assemble_attr) {
            NB2_CUDA(ctx, cudaFuncSetAttribute(k_mb_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin));
            S->assemble_attr = true;
        }
        k_mb_assemble<<<(n_mb + wpb - 1) / wpb, 32 * wpb, stage ? (size_t)wpb * words * 4 : 0, ctx->stream>>>(
            V, C, R, K, S->mb_of_link.p, ctx->params.warmstart_coeff, ctx->params.restitution_velocity_threshold, ctx->inv_dt,
            stage ? words : 0u, S->jac_words, ctx->contact_model);
    }
    k_mb_velocity_solve<<<(n_mb + MB_WPB - 1) / MB_WPB, 32 * MB_WPB, 0, ctx->stream>>>(V, R, K, C, ctx->contacts.p,
                                                                                     (int)ctx->params.max_velocity_iterations, ctx->params.dt);
    ctx->launches += 4;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

// step stage 5: position resolution; then the kinematics and dynamics of the end of the step (mechanical_world.rs:343-346)
int mb_launch_position(Context* ctx) {
    MbState* S = mb_state(ctx);
    if (!S || S->n_mb == 0) return NB2_OK;
    PosParams P;
    P.erp = ctx->params.erp;
    P.allowed_lin = ctx->params.allowed_linear_error;
    P.allowed_ang = ctx->params.allowed_angular_error;
    P.max_lin = ctx->params.max_linear_correction;
    P.max_ang = ctx->params.max_angular_correction;
    MbView V = mb_view(S);
    if (ctx->params.max_position_iterations > 0)
        k_mb_position_solve<<<mb_blocks(S->n_mb), MB_TPB, 0, ctx->stream>>>(V, mb_proxies(ctx), mb_contacts(ctx, S), mb_rows(S), P,
                                                                           (int)ctx->params.max_position_iterations);
    ctx->launches += 1;
    NB2_CUDA(ctx, cudaGetLastError());
    return mb_refresh(ctx, S, 2);
}

}  // namespace nb2
