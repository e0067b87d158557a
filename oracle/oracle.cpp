/*
 * oracle.cpp -- CPU restatement of nphysics' MoreauJeanSolver step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (nphysics_b200/)
 * may include, link or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker
 * and as the timed CPU baseline ("kind": "port").
 *
 * PARITY UNPINNED: the reference (Rust, nphysics3d 0.23) cannot be compiled in
 * this environment (no rustc/cargo, nalgebra 0.28 / ncollide3d 0.31 not
 * vendored) and ships no test, fixture or golden vector for this path
 * (SURVEY.md section 4 / 8c).  This file therefore follows the cited reference
 * lines one by one, single-threaded, f32, sequential Gauss-Seidel in the
 * reference's exact row order, and is validated by physics invariants in
 * tests/test_oracle_*.py.  Third-party arithmetic (nalgebra / ncollide) is
 * restated from its published algorithms; each such function says so.
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (no FMA contraction,
 * so the summation order below is what executes).
 *
 * All `file:line` citations are relative to /root/reference/.
 */
#include "../include/nphysics_b200.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

#ifndef NBO_REAL
#define NBO_REAL float
#endif
typedef NBO_REAL real;

namespace {

const real REAL_MAX = std::numeric_limits<real>::max();
const real REAL_EPS = std::numeric_limits<real>::epsilon();
const real REAL_PI = (real)3.14159265358979323846;

/* ---------------------------------------------------------------- algebra */
struct V3 {
    real x, y, z;
};
inline V3 v3(real x, real y, real z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, real s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, real s) { return v3(a.x / s, a.y / s, a.z / s); }
/* nalgebra dot on Vector3: a0*b0 + a1*b1 + a2*b2 (blas.rs dotx, U3 fast path). */
inline real dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* nalgebra Vector3::cross (matrix/ops cross, 3D branch). */
inline V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline real norm_squared(V3 a) { return dot(a, a); }
inline real norm(V3 a) { return std::sqrt(norm_squared(a)); }
inline real get(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* Row-major 3x3. */
struct M3 {
    real m[3][3];
};
inline M3 m3_zero() {
    M3 r;
    std::memset(&r, 0, sizeof(r));
    return r;
}
inline V3 operator*(const M3& a, V3 v) {
    return v3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
              a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
              a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
inline M3 operator*(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
inline M3 transpose(const M3& a) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
inline M3 operator+(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
    return r;
}
inline M3 operator-(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
    return r;
}
/* Vector3::cross_matrix (src/utils/generalized_cross.rs:70-85 -> nalgebra). */
inline M3 cross_matrix(V3 v) {
    M3 r = m3_zero();
    r.m[0][1] = -v.z;
    r.m[0][2] = v.y;
    r.m[1][0] = v.z;
    r.m[1][2] = -v.x;
    r.m[2][0] = -v.y;
    r.m[2][1] = v.x;
    return r;
}
/* nalgebra Matrix3::try_inverse (linalg/inverse.rs, 3x3 closed form):
 * false iff the determinant is exactly zero. */
inline bool try_inverse(const M3& a, M3* out) {
    const real m11 = a.m[0][0], m12 = a.m[0][1], m13 = a.m[0][2];
    const real m21 = a.m[1][0], m22 = a.m[1][1], m23 = a.m[1][2];
    const real m31 = a.m[2][0], m32 = a.m[2][1], m33 = a.m[2][2];
    const real minor_m12_m23 = m22 * m33 - m32 * m23;
    const real minor_m11_m23 = m21 * m33 - m31 * m23;
    const real minor_m11_m22 = m21 * m32 - m31 * m22;
    const real det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (det == (real)0) return false;
    out->m[0][0] = minor_m12_m23 / det;
    out->m[0][1] = (m13 * m32 - m33 * m12) / det;
    out->m[0][2] = (m12 * m23 - m22 * m13) / det;
    out->m[1][0] = -minor_m11_m23 / det;
    out->m[1][1] = (m11 * m33 - m31 * m13) / det;
    out->m[1][2] = (m13 * m21 - m23 * m11) / det;
    out->m[2][0] = minor_m11_m22 / det;
    out->m[2][1] = (m12 * m31 - m32 * m11) / det;
    out->m[2][2] = (m11 * m22 - m21 * m12) / det;
    return true;
}

/* Unit quaternion, nalgebra storage order (i, j, k, w). */
struct Quat {
    real i, j, k, w;
};
inline Quat quat_identity() { return Quat{0, 0, 0, 1}; }
/* nalgebra Quaternion * Quaternion (geometry/quaternion_ops.rs); no renormalisation. */
inline Quat operator*(Quat a, Quat b) {
    Quat r;
    r.w = a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k;
    r.i = a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j;
    r.j = a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i;
    r.k = a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w;
    return r;
}
inline Quat conjugate(Quat q) { return Quat{-q.i, -q.j, -q.k, q.w}; }
inline V3 imag(Quat q) { return v3(q.i, q.j, q.k); }
/* UnitQuaternion * Vector3: t = 2 (v x p); p + w t + v x t  (quaternion_ops.rs). */
inline V3 rotate(Quat q, V3 p) {
    V3 t = cross(imag(q), p) * (real)2;
    V3 c = cross(imag(q), t);
    return t * q.w + c + p;
}
inline V3 inv_rotate(Quat q, V3 p) { return rotate(conjugate(q), p); }
/* UnitQuaternion::to_rotation_matrix (geometry/quaternion.rs). */
inline M3 to_rotation_matrix(Quat q) {
    const real i = q.i, j = q.j, k = q.k, w = q.w;
    const real ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    const real ij = i * j * (real)2, wk = w * k * (real)2, wj = w * j * (real)2;
    const real ik = i * k * (real)2, jk = j * k * (real)2, wi = w * i * (real)2;
    M3 r;
    r.m[0][0] = ww + ii - jj - kk;
    r.m[0][1] = ij - wk;
    r.m[0][2] = wj + ik;
    r.m[1][0] = wk + ij;
    r.m[1][1] = ww - ii + jj - kk;
    r.m[1][2] = jk - wi;
    r.m[2][0] = ik - wj;
    r.m[2][1] = wi + jk;
    r.m[2][2] = ww - ii - jj + kk;
    return r;
}
/* UnitQuaternion::from_scaled_axis = exp(quat(0, w/2)), Quaternion::exp_eps with
 * eps = f32::EPSILON: identity when |h|^2 <= eps^2. */
inline Quat from_scaled_axis(V3 axisangle) {
    V3 h = axisangle / (real)2;
    real nn = norm_squared(h);
    if (nn <= REAL_EPS * REAL_EPS) return quat_identity();
    real w_exp = (real)1; /* exp(0) */
    real n = std::sqrt(nn);
    V3 nv = h * (w_exp * std::sin(n) / n);
    return Quat{nv.x, nv.y, nv.z, w_exp * std::cos(n)};
}
/* UnitQuaternion::from_axis_angle. */
inline Quat from_axis_angle(V3 axis, real angle) {
    real s = std::sin(angle / (real)2), c = std::cos(angle / (real)2);
    V3 v = axis * s;
    return Quat{v.x, v.y, v.z, c};
}
/* Unit::try_new_and_get(v, eps): Some((v/|v|, |v|)) iff |v|^2 > eps^2. */
inline bool try_new_and_get(V3 v, real eps, V3* dir, real* len) {
    real sq = norm_squared(v);
    if (sq > eps * eps) {
        real n = std::sqrt(sq);
        *dir = v / n;
        *len = n;
        return true;
    }
    return false;
}
/* UnitQuaternion::scaled_axis: axis() * angle(), zero when the imaginary part
 * is exactly zero; angle = 2 atan2(|v|, |w|), axis sign follows w. */
inline V3 scaled_axis(Quat q) {
    V3 v = q.w >= (real)0 ? imag(q) : -imag(q);
    real n = norm(v);
    if (n == (real)0) return v3(0, 0, 0);
    V3 axis = v / n;
    real angle = std::atan2(norm(imag(q)), std::fabs(q.w)) * (real)2;
    return axis * angle;
}
/* UnitQuaternion::rotation_between_axis(a, b) (scaled_rotation_between_axis, s=1). */
inline bool rotation_between_axis(V3 na, V3 nb, Quat* out) {
    V3 c = cross(na, nb);
    V3 axis;
    real len;
    if (try_new_and_get(c, REAL_EPS, &axis, &len)) {
        real cs = dot(na, nb);
        if (cs <= (real)-1) return false;
        if (cs >= (real)1) {
            *out = quat_identity();
            return true;
        }
        *out = from_axis_angle(axis, std::acos(cs));
        return true;
    } else if (dot(na, nb) < (real)0) {
        return false;
    }
    *out = quat_identity();
    return true;
}
/* Vector3::orthonormal_subspace_basis(&[n], f): calls f(a x n) then f(a)
 * (nalgebra geometry/... FiniteDimInnerSpace for U3). */
inline void orthonormal_subspace_basis(V3 n, V3* t1, V3* t2) {
    V3 a;
    if (std::fabs(n.x) > std::fabs(n.y))
        a = v3(n.z, 0, -n.x);
    else
        a = v3(0, -n.z, n.y);
    a = a / norm(a);
    *t1 = cross(a, n);
    *t2 = a;
}

struct Iso {
    V3 t;
    Quat r;
};
/* Isometry * Isometry: (R1 R2, t1 + R1 t2). */
inline Iso operator*(const Iso& a, const Iso& b) { return Iso{a.t + rotate(a.r, b.t), a.r * b.r}; }
inline V3 transform_point(const Iso& a, V3 p) { return rotate(a.r, p) + a.t; }

/* 6-vector, linear then angular (velocity3.rs:9-16, force3.rs:9-14). */
struct S6 {
    V3 lin, ang;
};
inline real at(const S6& s, int k) { return k < 3 ? get(s.lin, k) : get(s.ang, k - 3); }
/* nalgebra dot for a 6-vector slice: sequential res += a[k]*b[k] (blas.rs dotx
 * generic path with nrows < 8). */
inline real dot6(const real* a, const real* b) {
    real res = 0;
    for (int k = 0; k < 6; ++k) res += a[k] * b[k];
    return res;
}
/* axpy(a, x, 1): y[k] = a*x[k] + y[k]. */
inline void axpy6(real a, const real* x, real* y) {
    for (int k = 0; k < 6; ++k) y[k] = a * x[k] + y[k];
}

inline real dotn(size_t n, const real* a, const real* b);
inline void axpyn(size_t n, real a, const real* x, real* y);

/* Inertia3 (src/algebra/inertia3.rs:8-13). */
struct Inertia {
    real linear;
    M3 angular;
};
/* Inertia3::inverse (inertia3.rs:76-86). */
inline Inertia inertia_inverse(const Inertia& in) {
    Inertia r;
    r.linear = in.linear == (real)0 ? (real)0 : (real)1 / in.linear;
    if (!try_inverse(in.angular, &r.angular)) r.angular = m3_zero();
    return r;
}

/* ------------------------------------------------------------------ bodies */
struct Multibody;
struct Body {
    /* inputs (rigid_body.rs:26-50) */
    Iso position;
    S6 velocity;
    V3 local_com;
    Inertia local_inertia;
    S6 external_forces;
    real linear_damping, angular_damping, max_linear_velocity, max_angular_velocity;
    real jacobian_mask[6];
    uint32_t status;
    bool gravity_enabled;
    /* derived */
    V3 com;
    Inertia inertia, augmented_mass, inv_augmented_mass;
    S6 acceleration;
    size_t companion_id;
    /* ActivationStatus (body.rs:65-125): threshold < 0 stands for None (the body never sleeps) */
    real act_threshold = -1, act_energy = (real)0.04;
    bool is_active() const { return act_energy != 0; } /* body.rs:93-96 */

    /* a NB2_BODY_MULTIBODY_LINK record: the link it stands for (multibody.inc) */
    Multibody* mb = nullptr;
    int mb_link = -1;
    int handle = -1; /* the BodyHandle: the body's own index, or one value per multibody for its links */

    size_t status_dependent_ndofs() const; /* body.rs:287-293 */

    /* rigid_body.rs:305-315 (no renormalisation: improved_fixed_point_support off). */
    void set_position(const Iso& pos) {
        position = pos;
        com = transform_point(pos, local_com);
    }
    /* rigid_body.rs:371-381 + velocity3.rs:57-59 + Isometry3::new. */
    void apply_displacement(const real* d) {
        V3 lin = v3(d[0], d[1], d[2]), ang = v3(d[3], d[4], d[5]);
        Iso disp{lin, from_scaled_axis(ang)};
        /* shift * disp * shift.inverse(): rotation R, translation (com + t) + R(-com) */
        Iso wrt_com{(com + disp.t) + rotate(disp.r, -com), disp.r};
        set_position(wrt_com * position);
    }
    /* rigid_body.rs:467-505. */
    void integrate(real dt) {
        velocity.lin = velocity.lin * ((real)1 / ((real)1 + dt * linear_damping));
        velocity.ang = velocity.ang * ((real)1 / ((real)1 + dt * angular_damping));
        real linvel_norm = norm(velocity.lin);
        if (linvel_norm > max_linear_velocity) {
            if (max_linear_velocity == (real)0)
                velocity.lin = v3(0, 0, 0);
            else
                velocity.lin = velocity.lin * (max_linear_velocity / linvel_norm);
        }
        real angvel_norm = norm(velocity.ang);
        if (angvel_norm > max_angular_velocity) {
            if (max_angular_velocity == (real)0)
                velocity.ang = v3(0, 0, 0);
            else
                velocity.ang = velocity.ang * (max_angular_velocity / angvel_norm);
        }
        V3 dl = velocity.lin * dt, da = velocity.ang * dt;
        real disp[6] = {dl.x, dl.y, dl.z, da.x, da.y, da.z};
        apply_displacement(disp);
    }
    /* rigid_body.rs:558-588 (+ inertia3.rs:68-71). */
    void update_dynamics(real dt) {
        if (status != NB2_BODY_DYNAMIC) return;
        M3 rot = to_rotation_matrix(position.r);
        inertia.linear = local_inertia.linear;
        inertia.angular = (rot * local_inertia.angular) * transpose(rot);
        augmented_mass = inertia;
        const M3& i = inertia.angular;
        V3 w = velocity.ang;
        V3 iw = i * w;
        V3 w_dt = w * dt;
        M3 w_dt_cross = cross_matrix(w_dt);
        M3 iw_dt_cross = cross_matrix(iw * dt);
        augmented_mass.angular = augmented_mass.angular + (w_dt_cross * i - iw_dt_cross);
        inv_augmented_mass = inertia_inverse(augmented_mass);
    }
    /* rigid_body.rs:590-619. */
    void update_acceleration(V3 gravity) {
        acceleration = S6{v3(0, 0, 0), v3(0, 0, 0)};
        if (status != NB2_BODY_DYNAMIC) return;
        V3 w = velocity.ang;
        V3 iw = inertia.angular * w;
        V3 gyroscopic = -cross(w, iw);
        acceleration.ang = inv_augmented_mass.angular * gyroscopic;
        if (inv_augmented_mass.linear != (real)0 && gravity_enabled) acceleration.lin = gravity;
        /* acceleration += inv_augmented_mass * external_forces (inertia3.rs:166-173) */
        acceleration.lin = acceleration.lin + external_forces.lin * inv_augmented_mass.linear;
        acceleration.ang = acceleration.ang + inv_augmented_mass.angular * external_forces.ang;
        acceleration.lin = v3(acceleration.lin.x * jacobian_mask[0], acceleration.lin.y * jacobian_mask[1],
                              acceleration.lin.z * jacobian_mask[2]);
        acceleration.ang = v3(acceleration.ang.x * jacobian_mask[3], acceleration.ang.y * jacobian_mask[4],
                              acceleration.ang.z * jacobian_mask[5]);
    }
};

/* helper.rs:17-33. */
struct ForceDirection {
    bool angular;
    V3 dir;
};
inline ForceDirection fd_linear(V3 d) { return ForceDirection{false, d}; }
inline ForceDirection fd_angular(V3 d) { return ForceDirection{true, d}; }
inline ForceDirection fd_neg(const ForceDirection& f) { return ForceDirection{f.angular, -f.dir}; }

/* RigidBody::fill_constraint_geometry, rigid_body.rs:672-722. */
void mb_fill_constraint_geometry(const Body& b, V3 point, const ForceDirection& fdir, size_t j_id, size_t wj_id,
                                 real* jacobians, real* inv_r, const real* ext_vels, real* out_vel);
void fill_constraint_geometry(const Body& b, V3 point, const ForceDirection& fdir, size_t j_id, size_t wj_id,
                              real* jacobians, real* inv_r, const real* ext_vels, real* out_vel) {
    if (b.status == NB2_BODY_MULTIBODY_LINK) { /* Multibody::fill_constraint_geometry, multibody.rs:971-1025 */
        mb_fill_constraint_geometry(b, point, fdir, j_id, wj_id, jacobians, inv_r, ext_vels, out_vel);
        return;
    }
    V3 pos = point - b.com;
    /* ForceDirection::at_point (helper.rs:27-32), force3.rs:63-89 */
    S6 force = fdir.angular ? S6{v3(0, 0, 0), fdir.dir} : S6{fdir.dir, cross(pos, fdir.dir)};
    real f[6] = {force.lin.x, force.lin.y, force.lin.z, force.ang.x, force.ang.y, force.ang.z};
    real mf[6];
    for (int k = 0; k < 6; ++k) mf[k] = f[k] * b.jacobian_mask[k];
    real vel[6] = {b.velocity.lin.x, b.velocity.lin.y, b.velocity.lin.z,
                   b.velocity.ang.x, b.velocity.ang.y, b.velocity.ang.z};
    switch (b.status) {
        case NB2_BODY_KINEMATIC:
            if (out_vel) *out_vel += dot6(f, vel);
            break;
        case NB2_BODY_DYNAMIC: {
            for (int k = 0; k < 6; ++k) jacobians[j_id + k] = mf[k];
            const Inertia& im = b.inv_augmented_mass;
            V3 imf_lin = v3(mf[0], mf[1], mf[2]) * im.linear;
            V3 imf_ang = im.angular * v3(mf[3], mf[4], mf[5]);
            real imf[6] = {imf_lin.x, imf_lin.y, imf_lin.z, imf_ang.x, imf_ang.y, imf_ang.z};
            for (int k = 0; k < 6; ++k) jacobians[wj_id + k] = imf[k];
            *inv_r += im.linear + dot(v3(mf[3], mf[4], mf[5]), imf_ang);
            if (out_vel) {
                *out_vel += dot6(f, vel);
                if (ext_vels) *out_vel += dot6(mf, ext_vels);
            }
            break;
        }
        default:
            break;
    }
}

/* constraint.rs:5-37. */
struct ConstraintGeometry {
    size_t j_id1 = 0, j_id2 = 0, wj_id1 = 0, wj_id2 = 0, ndofs1 = 0, ndofs2 = 0;
    real r = 0;
    bool is_ground() const { return ndofs1 == 0 || ndofs2 == 0; }
};

/* helper::constraint_pair_geometry, helper.rs:53-135. */
ConstraintGeometry constraint_pair_geometry(const Body& body1, int h1, const Body& body2, int h2, V3 center1,
                                            V3 center2, const ForceDirection& dir, size_t* ground_j_id,
                                            size_t* j_id, std::vector<real>& jacobians, const real* ext_vels1,
                                            const real* ext_vels2, real* out_vel) {
    ConstraintGeometry res;
    res.ndofs1 = body1.status_dependent_ndofs();
    res.ndofs2 = body2.status_dependent_ndofs();
    size_t* out_j_id;
    if (res.ndofs1 == 0 || res.ndofs2 == 0) {
        res.j_id1 = *ground_j_id;
        out_j_id = ground_j_id;
    } else {
        res.j_id1 = *j_id;
        out_j_id = j_id;
    }
    res.j_id2 = res.j_id1 + res.ndofs1;
    res.wj_id1 = res.j_id2 + res.ndofs2;
    res.wj_id2 = res.wj_id1 + res.ndofs1;
    size_t need = res.wj_id2 + res.ndofs2;
    if (jacobians.size() < need) jacobians.resize(need, 0);
    real inv_r = 0;
    fill_constraint_geometry(body1, center1, dir, res.j_id1, res.wj_id1, jacobians.data(), &inv_r, ext_vels1,
                             out_vel);
    fill_constraint_geometry(body2, center2, fd_neg(dir), res.j_id2, res.wj_id2, jacobians.data(), &inv_r,
                             ext_vels2, out_vel);
    if (h1 == h2) { /* helper.rs:118-125: both parts on the same body */
        real c = 0;
        if (res.ndofs1 == res.ndofs2 && res.ndofs1 != 0)
            c = dotn(res.ndofs1, &jacobians[res.j_id2], &jacobians[res.wj_id1]) +
                dotn(res.ndofs1, &jacobians[res.j_id1], &jacobians[res.wj_id2]);
        inv_r += c;
    }
    res.r = inv_r != (real)0 ? (real)1 / inv_r : (real)1;
    *out_j_id += (res.ndofs1 + res.ndofs2) * 2;
    return res;
}

/* -------------------------------------------------------- constraint rows */
enum LimitKind { LIMIT_INDEPENDENT, LIMIT_DEPENDENT };
struct ImpulseLimits { /* constraint.rs:166-184 */
    LimitKind kind;
    real min, max;     /* Independent */
    size_t dependency; /* Dependent */
    real coeff;
};
struct Unilateral { /* constraint.rs:45-76 */
    real impulse, r, rhs;
    uint64_t impulse_id;
    size_t assembly_id1, assembly_id2, j_id1, j_id2, wj_id1, wj_id2, ndofs1, ndofs2;
};
struct UnilateralGround { /* constraint.rs:108-164 */
    real impulse, r, rhs;
    uint64_t impulse_id;
    size_t assembly_id, j_id, wj_id, ndofs;
};
struct Bilateral { /* constraint.rs:186-220 */
    real impulse, r, rhs;
    ImpulseLimits limits;
    uint64_t impulse_id;
    size_t assembly_id1, assembly_id2, j_id1, j_id2, wj_id1, wj_id2, ndofs1, ndofs2;
};
struct BilateralGround { /* constraint.rs:252-316 */
    real impulse, r, rhs;
    ImpulseLimits limits;
    uint64_t impulse_id;
    size_t assembly_id, j_id, wj_id, ndofs;
};
struct LinearConstraints { /* constraint_set.rs:9-50 */
    std::vector<Unilateral> unilateral;
    std::vector<UnilateralGround> unilateral_ground;
    std::vector<Bilateral> bilateral;
    std::vector<BilateralGround> bilateral_ground;
    void clear() {
        unilateral.clear();
        unilateral_ground.clear();
        bilateral.clear();
        bilateral_ground.clear();
    }
    size_t len() const {
        return unilateral.size() + unilateral_ground.size() + bilateral.size() + bilateral_ground.size();
    }
};
Unilateral make_unilateral(const ConstraintGeometry& g, size_t a1, size_t a2, real rhs, real impulse, uint64_t id) {
    return Unilateral{impulse, g.r, rhs, id, a1, a2, g.j_id1, g.j_id2, g.wj_id1, g.wj_id2, g.ndofs1, g.ndofs2};
}
UnilateralGround make_unilateral_ground(const ConstraintGeometry& g, size_t a1, size_t a2, real rhs, real impulse,
                                        uint64_t id) {
    if (g.ndofs1 == 0) return UnilateralGround{impulse, g.r, rhs, id, a2, g.j_id2, g.wj_id2, g.ndofs2};
    return UnilateralGround{impulse, g.r, rhs, id, a1, g.j_id1, g.wj_id1, g.ndofs1};
}
Bilateral make_bilateral(const ConstraintGeometry& g, size_t a1, size_t a2, ImpulseLimits lim, real rhs, real impulse,
                         uint64_t id) {
    return Bilateral{impulse, g.r, rhs, lim, id, a1, a2, g.j_id1, g.j_id2, g.wj_id1, g.wj_id2, g.ndofs1, g.ndofs2};
}
BilateralGround make_bilateral_ground(const ConstraintGeometry& g, size_t a1, size_t a2, ImpulseLimits lim, real rhs,
                                      real impulse, uint64_t id) {
    if (g.ndofs1 == 0) return BilateralGround{impulse, g.r, rhs, lim, id, a2, g.j_id2, g.wj_id2, g.ndofs2};
    return BilateralGround{impulse, g.r, rhs, lim, id, a1, g.j_id1, g.wj_id1, g.ndofs1};
}

#include "multibody.inc"

size_t Body::status_dependent_ndofs() const {
    if (status == NB2_BODY_MULTIBODY_LINK) return mb ? mb->ndofs : 0;
    return status == NB2_BODY_DYNAMIC ? 6 : 0;
}
void mb_fill_constraint_geometry(const Body& b, V3 point, const ForceDirection& fdir, size_t j_id, size_t wj_id,
                                 real* jacobians, real* inv_r, const real* ext_vels, real* out_vel) {
    b.mb->fill_constraint_geometry(b.mb_link, point, fdir, j_id, wj_id, jacobians, inv_r, ext_vels, out_vel);
}

/* nonlinear_constraint.rs:60-139: the contact position constraint. */
struct NonlinearUnilateral {
    int body1, body2;
    size_t ndofs1, ndofs2;
    V3 normal1, normal2;
    /* the cloned + dilated ContactKinematic */
    V3 local1, local2, dir1, dir2;
    real margin1, margin2;
    uint8_t geom1, geom2;
    Iso coll1_wrt_body, coll2_wrt_body;
    real rhs, r;
};
/* nonlinear_constraint.rs:9-58. */
struct GenericNonlinear {
    int body1, body2;
    bool is_angular;
    size_t dim1, dim2, wj_id1, wj_id2;
    real rhs, r;
};

struct Contact {
    V3 world1, world2, normal;
    real depth;
};

/* ncollide ContactKinematic::contact restated (from the
 * published algorithm; only Plane/Point, Point/Plane, Point/Point occur for
 * the supported shapes.  Point/Point assumes neither shape's tangent cone
 * contains the separation direction (true for balls). */
bool kinematic_contact(const NonlinearUnilateral& c, const Iso& m1, const Iso& m2, Contact* out) {
    V3 world1 = transform_point(m1, c.local1);
    V3 world2 = transform_point(m2, c.local2);
    V3 normal;
    real depth;
    if (c.geom1 == NB2_GEOM_PLANE && c.geom2 == NB2_GEOM_POINT) {
        normal = rotate(m1.r, c.dir1);
        depth = -dot(normal, world2 - world1);
        world1 = world2 + normal * depth;
    } else if (c.geom1 == NB2_GEOM_POINT && c.geom2 == NB2_GEOM_PLANE) {
        V3 world_normal2 = rotate(m2.r, c.dir2);
        depth = -dot(world_normal2, world1 - world2);
        world2 = world1 + world_normal2 * depth;
        normal = -world_normal2;
    } else if (c.geom1 != NB2_GEOM_PLANE && c.geom2 != NB2_GEOM_PLANE && c.geom1 <= NB2_GEOM_PLANE &&
               c.geom2 <= NB2_GEOM_PLANE) {
        /* Point/Point, Line/Line, Line/Point, Point/Line: a Line side is first
         * reduced to its point closest to the other side -- closest points of the two world lines, or the
         * projection of the other side's point on the line -- then the pair is resolved as Point/Point.
         * ncollide additionally asks shape 1 whether its tangent cone at the feature contains the
         * direction (the features interpenetrate: depth = +len, normal = -dir); the shapes are not part
         * of the contact record, so the separated branch is taken: with ncollide's margins the
         * un-dilated cores the features belong to do not interpenetrate in a resting contact. */
        if (c.geom1 == NB2_GEOM_LINE && c.geom2 == NB2_GEOM_LINE) {
            const V3 d1 = rotate(m1.r, c.dir1), d2 = rotate(m2.r, c.dir2);
            const V3 r = world1 - world2;
            const real a = dot(d1, d1), b = dot(d1, d2), cc = dot(d2, d2), d = dot(d1, r), e = dot(d2, r);
            const real denom = a * cc - b * b;
            real s1, s2;
            if (denom <= REAL_EPS * a * cc) { /* parallel lines: keep point 1, project it on line 2 */
                s1 = 0;
                s2 = cc != 0 ? e / cc : 0;
            } else {
                s1 = (b * e - cc * d) / denom;
                s2 = (a * e - b * d) / denom;
            }
            world1 = world1 + d1 * s1;
            world2 = world2 + d2 * s2;
        } else if (c.geom1 == NB2_GEOM_LINE) {
            const V3 d1 = rotate(m1.r, c.dir1);
            const real a = dot(d1, d1);
            if (a != 0) world1 = world1 + d1 * (dot(d1, world2 - world1) / a);
        } else if (c.geom2 == NB2_GEOM_LINE) {
            const V3 d2 = rotate(m2.r, c.dir2);
            const real a = dot(d2, d2);
            if (a != 0) world2 = world2 + d2 * (dot(d2, world1 - world2) / a);
        }
        V3 n;
        real d;
        if (try_new_and_get(world2 - world1, REAL_EPS, &n, &d)) {
            depth = -d;
            normal = n;
        } else {
            depth = 0;
            normal = rotate(m1.r, c.normal1);
        }
    } else {
        return false; /* Plane/Plane, Plane/Line, Line/Plane: ContactKinematic::contact returns None */
    }
    world1 = world1 + normal * c.margin1;
    world2 = world2 + normal * (-c.margin2);
    depth += c.margin1 + c.margin2;
    *out = Contact{world1, world2, normal, depth};
    return true;
}

/* ------------------------------------------------------------------ joints */
struct Joint {
    nb2_joint rec;
    /* ranges of the rows this joint pushed (ball_constraint.rs:22-23 etc.) */
    size_t bg_first = 0, bg_last = 0, b_first = 0, b_last = 0;
};

inline V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
inline Quat ldq(const float* p) { return Quat{p[0], p[1], p[2], p[3]}; }

struct World {
    nb2_params params;
    real inv_dt;
    std::vector<Body> bodies;
    std::vector<nb2_manifold> uploaded_manifolds; /* every contact pair of the narrow phase */
    std::vector<nb2_manifold> manifolds;          /* the step's list (mechanical_world.rs:287-300) */
    std::vector<nb2_contact> contacts;
    bool sleeping_enabled = false;
    std::vector<Joint> joints;
    std::vector<Multibody> mbs;            /* reduced-coordinate bodies (multibody.inc) */
    std::vector<nb2_mb_link> mb_link_recs; /* as uploaded, for the download */
    void apply_body_displacement(int body, const real* d) { /* Body::apply_displacement of the body behind a part */
        Body& b = bodies[body];
        if (b.status == NB2_BODY_MULTIBODY_LINK) b.mb->apply_displacement(d, bodies);
        else b.apply_displacement(d);
    }

    /* MoreauJeanSolver state (moreau_jean_solver.rs:14-23) */
    std::vector<real> jacobians, mj_lambda_vel, ext_vels;
    LinearConstraints contact_vel, joint_vel;
    std::vector<NonlinearUnilateral> contact_pos;
    std::vector<int> island;
    std::vector<size_t> active_joints; /* mechanical_world.rs:274-279 */

    /* SignoriniCoulombPyramidModel state (signorini_coulomb_pyramid_model.rs:19-25) */
    std::unordered_map<uint64_t, V3> impulses;
    std::vector<V3> contact_impulses_out; /* per contact, upload order */
    /* maps a contact's rows back to the contact index for the download */
    std::vector<size_t> uni_contact, unig_contact, bil_contact, bilg_contact;
    std::vector<char> solved_scratch;

    nb2_stats stats;
    char last_error[256];

    World() {
        nb2_params p;
        std::memset(&p, 0, sizeof(p));
        params = p;
        inv_dt = 0;
        std::memset(&stats, 0, sizeof(stats));
        last_error[0] = 0;
    }

    /* ---------------------------------------------------------- joint rows */
    /* helper.rs:167-242 cancel_relative_linear_velocity_wrt_axis and
     * helper.rs:422-497 cancel_relative_angular_velocity_wrt_axis share this
     * body; `lo` is -MAX except for one-sided limits (unit_constraint.rs:90-100). */
    void push_joint_row(const Joint& j, V3 anchor1, V3 anchor2, const ForceDirection& force, real impulse,
                        uint64_t impulse_id, real lo, size_t* ground_j_id, size_t* j_id) {
        const Body& b1 = bodies[j.rec.body1];
        const Body& b2 = bodies[j.rec.body2];
        ImpulseLimits limits{LIMIT_INDEPENDENT, lo, REAL_MAX, 0, 0};
        real rhs = 0;
        const real* ev1 = b1.status_dependent_ndofs() ? &ext_vels[b1.companion_id] : nullptr;
        const real* ev2 = b2.status_dependent_ndofs() ? &ext_vels[b2.companion_id] : nullptr;
        ConstraintGeometry geom = constraint_pair_geometry(b1, j.rec.body1, b2, j.rec.body2, anchor1, anchor2, force,
                                                           ground_j_id, j_id, jacobians, ev1, ev2, &rhs);
        if (geom.ndofs1 == 0 || geom.ndofs2 == 0)
            joint_vel.bilateral_ground.push_back(
                make_bilateral_ground(geom, b1.companion_id, b2.companion_id, limits, rhs, impulse, impulse_id));
        else
            joint_vel.bilateral.push_back(
                make_bilateral(geom, b1.companion_id, b2.companion_id, limits, rhs, impulse, impulse_id));
    }
    /* helper.rs:247-293. */
    void cancel_relative_linear_velocity(const Joint& j, V3 a1, V3 a2, const float* impulses, uint64_t impulse_id,
                                         size_t* gj, size_t* jj) {
        const V3 basis[3] = {v3(1, 0, 0), v3(0, 1, 0), v3(0, 0, 1)};
        for (int i = 0; i < 3; ++i)
            push_joint_row(j, a1, a2, fd_linear(basis[i]), impulses[i], impulse_id + i, -REAL_MAX, gj, jj);
    }
    /* helper.rs:502-548. */
    void cancel_relative_angular_velocity(const Joint& j, V3 a1, V3 a2, const float* impulses, uint64_t impulse_id,
                                          size_t* gj, size_t* jj) {
        const V3 basis[3] = {v3(1, 0, 0), v3(0, 1, 0), v3(0, 0, 1)};
        for (int i = 0; i < 3; ++i)
            push_joint_row(j, a1, a2, fd_angular(basis[i]), impulses[i], impulse_id + i, -REAL_MAX, gj, jj);
    }
    /* helper.rs:614-695. */
    void restrict_relative_angular_velocity_to_axis(const Joint& j, V3 axis, V3 a1, V3 a2, const float* impulses,
                                                    uint64_t impulse_id, size_t* gj, size_t* jj) {
        V3 t[2];
        orthonormal_subspace_basis(axis, &t[0], &t[1]);
        for (int i = 0; i < 2; ++i)
            push_joint_row(j, a1, a2, fd_angular(t[i]), impulses[i], impulse_id + i, -REAL_MAX, gj, jj);
    }
    /* helper.rs:771-853. */
    void restrict_relative_linear_velocity_to_axis(const Joint& j, V3 a1, V3 a2, V3 axis, const float* impulses,
                                                   uint64_t impulse_id, size_t* gj, size_t* jj) {
        V3 t[2];
        orthonormal_subspace_basis(axis, &t[0], &t[1]);
        for (int i = 0; i < 2; ++i)
            push_joint_row(j, a1, a2, fd_linear(t[i]), impulses[i], impulse_id + i, -REAL_MAX, gj, jj);
    }
    /* unit_constraint.rs:10-125. */
    void build_linear_limits_velocity_constraint(const Joint& j, V3 a1, V3 a2, V3 axis, bool has_min, real mn,
                                                 bool has_max, real mx, real impulse, uint64_t impulse_id, size_t* gj,
                                                 size_t* jj) {
        real offset = dot(axis, a2 - a1);
        bool unilateral;
        V3 dir;
        if (!has_min && !has_max) return;
        if (has_min && has_max) {
            /* relative_eq!(min, max): approx's default f32 epsilon/max_relative = f32::EPSILON */
            real diff = std::fabs(mn - mx);
            real largest = std::max(std::fabs(mn), std::fabs(mx));
            bool eq = (mn == mx) || diff <= REAL_EPS || diff <= largest * REAL_EPS;
            if (eq) {
                unilateral = false;
                dir = axis;
            } else if (offset <= mn) {
                unilateral = true;
                dir = -axis;
            } else if (offset >= mx) {
                unilateral = true;
                dir = axis;
            } else
                return;
        } else if (has_min) {
            if (offset <= mn) {
                unilateral = true;
                dir = -axis;
            } else
                return;
        } else {
            if (offset >= mx) {
                unilateral = true;
                dir = axis;
            } else
                return;
        }
        push_joint_row(j, a1, a2, fd_linear(dir), impulse, impulse_id, unilateral ? (real)0 : -REAL_MAX, gj, jj);
    }

    /* position_at_material_point (rigid_body.rs:645-648): position * Translation(point). */
    static Iso position_at_material_point(const Body& b, V3 p) {
        return Iso{b.position.t + rotate(b.position.r, p), b.position.r};
    }

    /* Each *_constraint.rs::velocity_constraints. */
    void joint_velocity_constraints(Joint& j, size_t* gj, size_t* jj) {
        const nb2_joint& r = j.rec;
        const Body& body1 = bodies[r.body1];
        const Body& body2 = bodies[r.body2];
        j.bg_first = joint_vel.bilateral_ground.size();
        j.b_first = joint_vel.bilateral.size();
        Iso pos1 = position_at_material_point(body1, ld3(r.anchor1));
        Iso pos2 = position_at_material_point(body2, ld3(r.anchor2));
        if (r.type == NB2_JOINT_FIXED || r.type == NB2_JOINT_CARTESIAN) { /* fixed_constraint.rs:118-119 */
            pos1.r = pos1.r * ldq(r.ref_frame1);
            pos2.r = pos2.r * ldq(r.ref_frame2);
        }
        V3 anchor1 = pos1.t, anchor2 = pos2.t;
        const float* lin = &r.impulses[0];
        const float* ang = &r.impulses[3];
        switch (r.type) {
            case NB2_JOINT_BALL: /* ball_constraint.rs:92-126 */
                cancel_relative_linear_velocity(j, anchor1, anchor2, lin, 0, gj, jj);
                break;
            case NB2_JOINT_REVOLUTE: { /* revolute_constraint.rs:194-258 */
                cancel_relative_linear_velocity(j, anchor1, anchor2, lin, 0, gj, jj);
                V3 axis1 = rotate(pos1.r, ld3(r.axis1));
                restrict_relative_angular_velocity_to_axis(j, axis1, anchor1, anchor2, ang, 3, gj, jj);
                break;
            }
            case NB2_JOINT_PRISMATIC: { /* prismatic_constraint.rs:149-236 */
                V3 axis = rotate(pos1.r, ld3(r.axis1));
                restrict_relative_linear_velocity_to_axis(j, anchor1, anchor2, axis, lin, 0, gj, jj);
                cancel_relative_angular_velocity(j, anchor1, anchor2, ang, 2, gj, jj);
                build_linear_limits_velocity_constraint(j, anchor1, anchor2, axis, r.flags & NB2_JOINT_FLAG_MIN_OFFSET,
                                                        r.min_offset, r.flags & NB2_JOINT_FLAG_MAX_OFFSET,
                                                        r.max_offset, r.impulses[6], 5, gj, jj);
                break;
            }
            case NB2_JOINT_UNIVERSAL: { /* universal_constraint.rs:104-165 */
                cancel_relative_linear_velocity(j, anchor1, anchor2, lin, 0, gj, jj);
                V3 axis1 = rotate(pos1.r, ld3(r.axis1));
                V3 axis2 = rotate(pos2.r, ld3(r.axis2));
                V3 orth;
                real len;
                if (try_new_and_get(cross(axis1, axis2), REAL_EPS, &orth, &len))
                    push_joint_row(j, anchor1, anchor2, fd_angular(orth), r.impulses[3], 3, -REAL_MAX, gj, jj);
                break;
            }
            case NB2_JOINT_PLANAR: { /* planar_constraint.rs:101-163 */
                V3 axis1 = rotate(pos1.r, ld3(r.axis1));
                push_joint_row(j, anchor1, anchor2, fd_linear(axis1), r.impulses[0], 0, -REAL_MAX, gj, jj);
                restrict_relative_angular_velocity_to_axis(j, axis1, anchor1, anchor2, ang, 1, gj, jj);
                break;
            }
            case NB2_JOINT_RECTANGULAR: { /* rectangular_constraint.rs:99-160 */
                V3 axis1 = rotate(pos1.r, ld3(r.axis1));
                push_joint_row(j, anchor1, anchor2, fd_linear(axis1), r.impulses[0], 0, -REAL_MAX, gj, jj);
                cancel_relative_angular_velocity(j, anchor1, anchor2, ang, 1, gj, jj);
                break;
            }
            case NB2_JOINT_PIN_SLOT: { /* pin_slot_constraint.rs:149-212 */
                V3 axis_v1 = rotate(pos1.r, ld3(r.axis1));
                V3 axis_w1 = rotate(pos1.r, ld3(r.axis3));
                restrict_relative_linear_velocity_to_axis(j, anchor1, anchor2, axis_v1, lin, 0, gj, jj);
                restrict_relative_angular_velocity_to_axis(j, axis_w1, anchor1, anchor2, ang, 2, gj, jj);
                break;
            }
            case NB2_JOINT_CYLINDRICAL: { /* cylindrical_constraint.rs:143-205 */
                V3 axis1 = rotate(pos1.r, ld3(r.axis1));
                restrict_relative_linear_velocity_to_axis(j, anchor1, anchor2, axis1, lin, 0, gj, jj);
                restrict_relative_angular_velocity_to_axis(j, axis1, anchor1, anchor2, ang, 2, gj, jj);
                break;
            }
            case NB2_JOINT_FIXED: /* fixed_constraint.rs:113-171 */
                cancel_relative_linear_velocity(j, anchor1, anchor2, lin, 0, gj, jj);
                cancel_relative_angular_velocity(j, anchor1, anchor2, ang, 3, gj, jj);
                break;
            case NB2_JOINT_CARTESIAN: /* cartesian_constraint.rs:106-144 */
                cancel_relative_angular_velocity(j, anchor1, anchor2, ang, 0, gj, jj);
                break;
            default:
                break;
        }
        j.bg_last = joint_vel.bilateral_ground.size();
        j.b_last = joint_vel.bilateral.size();
    }

    /* Each cache_impulses, verbatim slot mapping (including the pin-slot /
     * cylindrical quirk where impulse_id 2 lands in lin_impulses[2]). */
    static void joint_store_impulse(nb2_joint& r, uint64_t id, real impulse) {
        float* lin = &r.impulses[0];
        float* ang = &r.impulses[3];
        switch (r.type) {
            case NB2_JOINT_BALL: lin[id] = impulse; break;          /* ball_constraint.rs:133-139 */
            case NB2_JOINT_REVOLUTE:                                  /* revolute_constraint.rs:267-281 */
            case NB2_JOINT_PIN_SLOT:                                  /* pin_slot_constraint.rs:222-236 */
            case NB2_JOINT_CYLINDRICAL:                               /* cylindrical_constraint.rs:215-229 */
            case NB2_JOINT_FIXED:                                     /* fixed_constraint.rs:181-195 */
                if (id < 3) lin[id] = impulse; else ang[id - 3] = impulse;
                break;
            case NB2_JOINT_PRISMATIC:                                 /* prismatic_constraint.rs:246-264 */
                if (id < 2) lin[id] = impulse;
                else if (id < 5) ang[id + 1 - 3] = impulse;
                else r.impulses[6] = impulse;
                break;
            case NB2_JOINT_UNIVERSAL:                                 /* universal_constraint.rs:175-189 */
                if (id < 3) lin[id] = impulse; else ang[0] = impulse;
                break;
            case NB2_JOINT_PLANAR:                                    /* planar_constraint.rs:173-187 */
            case NB2_JOINT_RECTANGULAR:                               /* rectangular_constraint.rs:170-184 */
                if (id == 0) lin[0] = impulse; else ang[id - 1] = impulse;
                break;
            case NB2_JOINT_CARTESIAN: ang[id] = impulse; break;      /* cartesian_constraint.rs:154-160 */
            default: break;
        }
    }
    void joint_cache_impulses(Joint& j) {
        nb2_joint& r = j.rec;
        for (size_t k = j.bg_first; k < j.bg_last; ++k)
            joint_store_impulse(r, joint_vel.bilateral_ground[k].impulse_id, joint_vel.bilateral_ground[k].impulse);
        for (size_t k = j.b_first; k < j.b_last; ++k)
            joint_store_impulse(r, joint_vel.bilateral[k].impulse_id, joint_vel.bilateral[k].impulse);
        real inv_dt2 = inv_dt * inv_dt;
        const float* lin = &r.impulses[0];
        const float* ang = &r.impulses[3];
        real lin_sq = norm_squared(v3(lin[0], lin[1], lin[2]));
        real ang_sq = norm_squared(v3(ang[0], ang[1], ang[2]));
        bool broken = false;
        switch (r.type) {
            case NB2_JOINT_BALL: /* ball_constraint.rs:141-143 */
                broken = lin_sq * inv_dt * inv_dt > r.break_force_squared;
                break;
            case NB2_JOINT_UNIVERSAL: /* universal_constraint.rs:191-197 */
                broken = lin_sq * inv_dt2 > r.break_force_squared ||
                         (real)ang[0] * ang[0] * inv_dt2 > r.break_torque_squared;
                break;
            case NB2_JOINT_PLANAR: /* planar_constraint.rs:189-197 */
                broken = (real)lin[0] * lin[0] * inv_dt2 > r.break_force_squared ||
                         (real)ang[0] * ang[0] * inv_dt2 + (real)ang[1] * ang[1] * inv_dt2 > r.break_torque_squared;
                break;
            case NB2_JOINT_RECTANGULAR: /* rectangular_constraint.rs:186-192 */
                broken = (real)lin[0] * lin[0] * inv_dt2 > r.break_force_squared ||
                         ang_sq * inv_dt2 > r.break_torque_squared;
                break;
            case NB2_JOINT_CARTESIAN: /* cartesian_constraint.rs:162-164 */
                broken = ang_sq * inv_dt * inv_dt > r.break_torque_squared;
                break;
            default: /* revolute :283-289, prismatic :266-272, pin-slot, cylindrical, fixed */
                broken = lin_sq * inv_dt2 > r.break_force_squared || ang_sq * inv_dt2 > r.break_torque_squared;
                break;
        }
        if (broken) r.broken = 1;
    }

    /* JointConstraint::is_active (joint_constraint.rs:219-228). */
    bool joint_is_active(const Joint& j) const {
        const Body& b1 = bodies[j.rec.body1];
        const Body& b2 = bodies[j.rec.body2];
        return (b1.status_dependent_ndofs() != 0 && b1.is_active()) || (b2.status_dependent_ndofs() != 0 && b2.is_active());
    }

    /* ------------------------------------------------------------ sleeping */
    /* utils/union_find.rs:30-59 */
    struct UnionFindSet {
        size_t parent, rank;
    };
    static size_t uf_find(size_t x, std::vector<UnionFindSet>& sets) {
        if (sets[x].parent != x) sets[x].parent = uf_find(sets[x].parent, sets);
        return sets[x].parent;
    }
    static void uf_union(size_t x, size_t y, std::vector<UnionFindSet>& sets) {
        size_t xr = uf_find(x, sets), yr = uf_find(y, sets);
        if (xr == yr) return;
        size_t rx = sets[xr].rank, ry = sets[yr].rank;
        if (rx < ry) sets[xr].parent = yr;
        else if (rx > ry) sets[yr].parent = xr;
        else {
            sets[yr].parent = xr;
            sets[xr].rank = rx + 1;
        }
    }
    /* ActivationManager::update (detection/activation_manager.rs:60-201).  `to_activate` are the
     * deferred_activate handles.  Contact pairs = every uploaded manifold with at least one contact. */
    void update_activation(real mix_factor, const int32_t* to_activate, uint32_t n_to_activate) {
        std::vector<size_t> id_to_body;
        std::vector<size_t> companion(bodies.size(), (size_t)-1);
        for (size_t i = 0; i < bodies.size(); ++i) { /* :79-93 */
            Body& b = bodies[i];
            if (b.status_dependent_ndofs() != 0) {
                if (b.is_active() && b.act_threshold >= 0) { /* update_energy :47-58 */
                    const real v[6] = {b.velocity.lin.x, b.velocity.lin.y, b.velocity.lin.z,
                                       b.velocity.ang.x, b.velocity.ang.y, b.velocity.ang.z};
                    real nsq = 0; /* nalgebra norm_squared of a 6-slice: sequential (appendix B) */
                    for (int k = 0; k < 6; ++k) nsq += v[k] * v[k];
                    real e = ((real)1 - mix_factor) * b.act_energy + mix_factor * nsq;
                    b.act_energy = std::min(e, b.act_threshold * (real)4);
                }
                companion[i] = id_to_body.size();
                id_to_body.push_back(i);
            }
            if (b.status == NB2_BODY_KINEMATIC) {
                companion[i] = id_to_body.size();
                id_to_body.push_back(i);
            }
        }
        for (uint32_t k = 0; k < n_to_activate; ++k) { /* :100-108; Body::activate body.rs:338-342 */
            if (to_activate[k] < 0 || (size_t)to_activate[k] >= bodies.size()) continue;
            Body& b = bodies[to_activate[k]];
            if (b.act_threshold >= 0) b.act_energy = b.act_threshold * (real)2;
        }
        std::vector<UnionFindSet> ufind(id_to_body.size());
        std::vector<char> can_deactivate(id_to_body.size(), 1);
        for (size_t i = 0; i < ufind.size(); ++i) ufind[i] = UnionFindSet{i, 0};
        auto make_union = [&](int h1, int h2) { /* :138-152 */
            const Body& b1 = bodies[h1];
            const Body& b2 = bodies[h2];
            if ((b1.status_dependent_ndofs() != 0 || b1.status == NB2_BODY_KINEMATIC) &&
                (b2.status_dependent_ndofs() != 0 || b2.status == NB2_BODY_KINEMATIC))
                uf_union(companion[h1], companion[h2], ufind);
        };
        for (const nb2_manifold& m : uploaded_manifolds)
            if (m.num_contacts > 0) make_union(m.body1, m.body2);
        for (const Joint& j : joints)
            if (!j.rec.broken) make_union(j.rec.body1, j.rec.body2);
        for (size_t i = 0; i < ufind.size(); ++i) { /* :170-182 */
            size_t root = uf_find(i, ufind);
            const Body& b = bodies[id_to_body[i]];
            can_deactivate[root] = b.act_threshold >= 0 ? (can_deactivate[root] && b.act_energy < b.act_threshold) : 0;
        }
        for (size_t i = 0; i < ufind.size(); ++i) { /* :185-206 */
            size_t root = uf_find(i, ufind);
            Body& b = bodies[id_to_body[i]];
            if (can_deactivate[root]) {
                if (b.is_active()) { /* RigidBody::deactivate rigid_body.rs:396-400 */
                    b.act_energy = 0;
                    b.velocity = S6{v3(0, 0, 0), v3(0, 0, 0)};
                }
            } else if (b.status != NB2_BODY_KINEMATIC) {
                if (!b.is_active() && b.act_threshold >= 0) b.act_energy = b.act_threshold * (real)2;
            }
        }
    }

    /* ----------------------------------------------- joint position rows */
    bool make_generic(const Joint& j, V3 a1, V3 a2, const ForceDirection& force, bool is_angular, real rhs,
                      GenericNonlinear* out) {
        size_t j_id = 0, ground_j_id = 0;
        ConstraintGeometry geom =
            constraint_pair_geometry(bodies[j.rec.body1], j.rec.body1, bodies[j.rec.body2], j.rec.body2, a1, a2, force,
                                     &ground_j_id, &j_id, pos_jacobians, nullptr, nullptr, nullptr);
        *out = GenericNonlinear{j.rec.body1, j.rec.body2, is_angular, geom.ndofs1, geom.ndofs2,
                                geom.wj_id1, geom.wj_id2, rhs, geom.r};
        return true;
    }
    /* helper.rs:298-359. */
    bool cancel_relative_translation_wrt_axis(const Joint& j, V3 a1, V3 a2, V3 axis, GenericNonlinear* out) {
        real depth = dot(axis, a2 - a1);
        ForceDirection force = fd_linear(axis);
        if (depth < (real)0) {
            depth = -depth;
            force = fd_linear(-axis);
        }
        if (depth > params.allowed_linear_error) return make_generic(j, a1, a2, force, false, -depth, out);
        return false;
    }
    /* helper.rs:364-417. */
    bool cancel_relative_translation(const Joint& j, V3 a1, V3 a2, GenericNonlinear* out) {
        V3 dir;
        real depth;
        if (try_new_and_get(a2 - a1, params.allowed_linear_error, &dir, &depth))
            return make_generic(j, a1, a2, fd_linear(dir), false, -depth, out);
        return false;
    }
    /* helper.rs:553-608. */
    bool cancel_relative_rotation(const Joint& j, V3 a1, V3 a2, Quat rot1, Quat rot2, GenericNonlinear* out) {
        V3 error = scaled_axis(rot2 * conjugate(rot1));
        V3 dir;
        real depth;
        if (try_new_and_get(error, params.allowed_angular_error, &dir, &depth))
            return make_generic(j, a1, a2, fd_angular(dir), true, -depth, out);
        return false;
    }
    static V3 pi_fallback_axis(V3 axis1) { /* helper.rs:720-724 */
        int imin = 0;
        real best = std::fabs(axis1.x);
        if (std::fabs(axis1.y) < best) {
            best = std::fabs(axis1.y);
            imin = 1;
        }
        if (std::fabs(axis1.z) < best) imin = 2;
        V3 e = v3(imin == 0, imin == 1, imin == 2);
        V3 c = cross(e, axis1);
        return (c / norm(c)) * REAL_PI;
    }
    /* helper.rs:701-766. */
    bool align_axis(const Joint& j, V3 a1, V3 a2, V3 axis1, V3 axis2, GenericNonlinear* out) {
        V3 error;
        Quat rot;
        if (rotation_between_axis(axis1, axis2, &rot))
            error = scaled_axis(rot);
        else
            error = pi_fallback_axis(axis1);
        V3 dir;
        real depth;
        if (try_new_and_get(error, params.allowed_angular_error, &dir, &depth))
            return make_generic(j, a1, a2, fd_angular(dir), true, -depth, out);
        return false;
    }
    /* helper.rs:858-915. */
    bool project_anchor_to_axis(const Joint& j, V3 a1, V3 a2, V3 axis1, GenericNonlinear* out) {
        V3 dpt = a2 - a1;
        V3 proj = a1 + axis1 * dot(axis1, dpt);
        V3 error = a2 - proj;
        V3 dir;
        real depth;
        if (try_new_and_get(error, params.allowed_linear_error, &dir, &depth))
            return make_generic(j, a1, a2, fd_linear(dir), false, -depth, out);
        return false;
    }
    /* helper.rs:921-998. */
    bool restore_angle_between_axis(const Joint& j, V3 a1, V3 a2, V3 axis1, V3 axis2, real angle,
                                    GenericNonlinear* out) {
        V3 separation;
        Quat rot;
        if (rotation_between_axis(axis1, axis2, &rot))
            separation = scaled_axis(rot);
        else
            separation = pi_fallback_axis(axis1);
        V3 dir;
        real curr_ang;
        if (try_new_and_get(separation, REAL_EPS, &dir, &curr_ang)) {
            real error = curr_ang - angle;
            if (error < (real)0) {
                error = -error;
                dir = -dir;
            }
            if (error < params.allowed_angular_error) return false;
            return make_generic(j, a1, a2, fd_angular(dir), true, -error, out);
        }
        return false;
    }
    /* unit_constraint.rs:127-197. */
    bool build_linear_limits_position_constraint(const Joint& j, V3 a1, V3 a2, V3 axis, bool has_min, real mn,
                                                 bool has_max, real mx, GenericNonlinear* out) {
        real offset = dot(axis, a2 - a1);
        real error = 0;
        V3 dir = axis;
        if (has_min) {
            error = mn - offset;
            dir = -axis;
        }
        if (error < (real)0) {
            if (has_max) {
                error = offset - mx;
                dir = axis;
            }
        }
        if (error > params.allowed_linear_error) return make_generic(j, a1, a2, fd_linear(dir), false, -error, out);
        return false;
    }

    size_t joint_num_position_constraints(const Joint& j) const {
        if (!joint_is_active(j)) return 0;
        switch (j.rec.type) {
            case NB2_JOINT_BALL: return 1;
            case NB2_JOINT_CARTESIAN: return 1;
            case NB2_JOINT_PRISMATIC:
                return (j.rec.flags & (NB2_JOINT_FLAG_MIN_OFFSET | NB2_JOINT_FLAG_MAX_OFFSET)) ? 3 : 2;
            default: return 2;
        }
    }
    bool joint_position_constraint(const Joint& j, size_t i, GenericNonlinear* out) {
        const nb2_joint& r = j.rec;
        const Body& body1 = bodies[r.body1];
        const Body& body2 = bodies[r.body2];
        Iso pos1 = position_at_material_point(body1, ld3(r.anchor1));
        Iso pos2 = position_at_material_point(body2, ld3(r.anchor2));
        if (r.type == NB2_JOINT_FIXED || r.type == NB2_JOINT_CARTESIAN) {
            pos1.r = pos1.r * ldq(r.ref_frame1);
            pos2.r = pos2.r * ldq(r.ref_frame2);
        }
        V3 a1 = pos1.t, a2 = pos2.t;
        switch (r.type) {
            case NB2_JOINT_BALL: /* ball_constraint.rs:158-177 */
                return cancel_relative_translation(j, a1, a2, out);
            case NB2_JOINT_REVOLUTE: /* revolute_constraint.rs:309-348 */
                if (i == 0) return cancel_relative_translation(j, a1, a2, out);
                if (i == 1)
                    return align_axis(j, a1, a2, rotate(pos1.r, ld3(r.axis1)), rotate(pos2.r, ld3(r.axis2)), out);
                return false;
            case NB2_JOINT_PRISMATIC: { /* prismatic_constraint.rs:288-347 */
                if (i == 0) return cancel_relative_rotation(j, a1, a2, pos1.r, pos2.r, out);
                V3 axis = rotate(pos1.r, ld3(r.axis1));
                if (i == 1) return project_anchor_to_axis(j, a1, a2, axis, out);
                if (i == 2)
                    return build_linear_limits_position_constraint(j, a1, a2, axis, r.flags & NB2_JOINT_FLAG_MIN_OFFSET,
                                                                   r.min_offset, r.flags & NB2_JOINT_FLAG_MAX_OFFSET,
                                                                   r.max_offset, out);
                return false;
            }
            case NB2_JOINT_UNIVERSAL: /* universal_constraint.rs:213-250 */
                if (i == 0) return cancel_relative_translation(j, a1, a2, out);
                if (i == 1)
                    return restore_angle_between_axis(j, a1, a2, rotate(pos1.r, ld3(r.axis1)),
                                                      rotate(pos2.r, ld3(r.axis2)), r.angle, out);
                return false;
            case NB2_JOINT_PLANAR: { /* planar_constraint.rs:213-249 */
                V3 axis1 = rotate(pos1.r, ld3(r.axis1));
                if (i == 0) return cancel_relative_translation_wrt_axis(j, a1, a2, axis1, out);
                if (i == 1) return align_axis(j, a1, a2, axis1, rotate(pos2.r, ld3(r.axis2)), out);
                return false;
            }
            case NB2_JOINT_RECTANGULAR: { /* rectangular_constraint.rs:208-252 */
                V3 axis1 = rotate(pos1.r, ld3(r.axis1));
                if (i == 0) return cancel_relative_translation_wrt_axis(j, a1, a2, axis1, out);
                if (i == 1) return cancel_relative_rotation(j, a1, a2, pos1.r, pos2.r, out);
                return false;
            }
            case NB2_JOINT_PIN_SLOT: { /* pin_slot_constraint.rs:262-297: anchors from part.position() */
                if (i == 0)
                    return align_axis(j, a1, a2, rotate(pos1.r, ld3(r.axis3)), rotate(pos2.r, ld3(r.axis2)), out);
                if (i == 1) return project_anchor_to_axis(j, a1, a2, rotate(pos1.r, ld3(r.axis1)), out);
                return false;
            }
            case NB2_JOINT_CYLINDRICAL: { /* cylindrical_constraint.rs:255-288 */
                V3 axis1 = rotate(pos1.r, ld3(r.axis1));
                if (i == 0) return align_axis(j, a1, a2, axis1, rotate(pos2.r, ld3(r.axis2)), out);
                if (i == 1) return project_anchor_to_axis(j, a1, a2, axis1, out);
                return false;
            }
            case NB2_JOINT_FIXED: /* fixed_constraint.rs:215-248 */
                if (i == 0) return cancel_relative_rotation(j, a1, a2, pos1.r, pos2.r, out);
                if (i == 1) return cancel_relative_translation(j, a1, a2, out);
                return false;
            case NB2_JOINT_CARTESIAN: /* cartesian_constraint.rs:180-200 */
                return cancel_relative_rotation(j, a1, a2, pos1.r, pos2.r, out);
            default:
                return false;
        }
    }

    /* --------------------------------------------------- contact assembly */
    /* SignoriniCoulombPyramidModel::constraints, signorini_coulomb_pyramid_model.rs:56-224
     * (+ SignoriniModel::build_velocity_constraint signorini_model.rs:37-138 and
     * build_position_constraint :153-197). */
    /* 0 = SignoriniCoulombPyramidModel (default, moreau_jean_solver.rs:29-40), 1 = SignoriniModel as a
     * frictionless ContactModel (signorini_model.rs:200-298), selected by set_contact_model (:42-44) */
    int contact_model = 0;

    void contact_constraints(size_t* ground_j_id, size_t* j_id) {
        uni_contact.clear();
        unig_contact.clear();
        bil_contact.clear();
        bilg_contact.clear();
        for (const nb2_manifold& m : manifolds) {
            const Body& body1 = bodies[m.body1];
            const Body& body2 = bodies[m.body2];
            for (uint32_t ci = m.first_contact; ci < m.first_contact + m.num_contacts; ++ci) {
                const nb2_contact& c = contacts[ci];
                /* SignoriniModel::is_constraint_active (signorini_model.rs:141-150, applied at :230-232;
                 * commented out in the pyramid model, signorini_coulomb_pyramid_model.rs:100-102) */
                if (contact_model == 1 && !((real)c.depth + (real)m.margin1 + (real)m.margin2 >= (real)0)) continue;
                V3 normal = ld3(c.normal), world1 = ld3(c.world1), world2 = ld3(c.world2);
                V3 surface_velocity = ld3(m.surface_velocity);
                V3 impulse = v3(0, 0, 0);
                if (c.key != 0) {
                    auto it = impulses.find(c.key);
                    if (it != impulses.end()) impulse = it->second;
                }
                size_t assembly_id1 = body1.companion_id, assembly_id2 = body2.companion_id;
                const real* ev1 = body1.status_dependent_ndofs() ? &ext_vels[assembly_id1] : nullptr;
                const real* ev2 = body2.status_dependent_ndofs() ? &ext_vels[assembly_id2] : nullptr;

                /* --- normal row: signorini_model.rs:65-137 */
                V3 center1 = world1 + normal * (real)m.margin1;
                V3 center2 = world2 - normal * (real)m.margin2;
                real rhs = dot(normal, surface_velocity);
                ConstraintGeometry geom =
                    constraint_pair_geometry(body1, body1.handle, body2, body2.handle, center1, center2, fd_linear(-normal),
                                             ground_j_id, j_id, jacobians, ev1, ev2, &rhs);
                if (rhs <= -params.restitution_velocity_threshold) rhs += (real)m.restitution * rhs;
                real depth = (real)c.depth + (real)m.margin1 + (real)m.margin2;
                if (depth < (real)0) rhs += (-depth) * inv_dt;
                real warmstart = impulse.x * params.warmstart_coeff;
                bool ground_constraint = geom.is_ground();
                if (ground_constraint) {
                    contact_vel.unilateral_ground.push_back(
                        make_unilateral_ground(geom, assembly_id1, assembly_id2, rhs, warmstart, c.key));
                    unig_contact.push_back(ci);
                } else {
                    contact_vel.unilateral.push_back(
                        make_unilateral(geom, assembly_id1, assembly_id2, rhs, warmstart, c.key));
                    uni_contact.push_back(ci);
                }

                /* --- position row: signorini_model.rs:153-197 */
                NonlinearUnilateral p;
                p.body1 = m.body1;
                p.body2 = m.body2;
                p.ndofs1 = body1.status_dependent_ndofs();
                p.ndofs2 = body2.status_dependent_ndofs();
                p.normal1 = inv_rotate(body1.position.r, normal);
                p.normal2 = -inv_rotate(body2.position.r, normal);
                p.local1 = ld3(c.local1);
                p.local2 = ld3(c.local2);
                p.dir1 = ld3(c.dir1);
                p.dir2 = ld3(c.dir2);
                p.margin1 = (real)c.dilation1 + (real)m.margin1;
                p.margin2 = (real)c.dilation2 + (real)m.margin2;
                p.geom1 = c.geom1;
                p.geom2 = c.geom2;
                p.coll1_wrt_body = Iso{ld3(m.coll1_wrt_body), ldq(m.coll1_wrt_body + 3)};
                p.coll2_wrt_body = Iso{ld3(m.coll2_wrt_body), ldq(m.coll2_wrt_body + 3)};
                p.rhs = 0;
                p.r = 0;
                contact_pos.push_back(p);

                if (contact_model == 1) continue; /* frictionless: the normal row and the position row only */
                /* --- friction rows: signorini_coulomb_pyramid_model.rs:131-216 */
                size_t dependency =
                    ground_constraint ? contact_vel.unilateral_ground.size() - 1 : contact_vel.unilateral.size() - 1;
                ImpulseLimits limits{LIMIT_DEPENDENT, 0, 0, dependency, (real)m.friction};
                V3 t[2];
                orthonormal_subspace_basis(normal, &t[0], &t[1]);
                for (int i = 0; i < 2; ++i) {
                    real frhs = dot(t[i], surface_velocity);
                    ConstraintGeometry fgeom =
                        constraint_pair_geometry(body1, body1.handle, body2, body2.handle, center1, center2, fd_linear(t[i]),
                                                 ground_j_id, j_id, jacobians, ev1, ev2, &frhs);
                    real fwarm = get(impulse, i + 1) * params.warmstart_coeff;
                    if (fgeom.is_ground()) {
                        contact_vel.bilateral_ground.push_back(
                            make_bilateral_ground(fgeom, assembly_id1, assembly_id2, limits, frhs, fwarm, c.key));
                        bilg_contact.push_back(ci);
                    } else {
                        contact_vel.bilateral.push_back(
                            make_bilateral(fgeom, assembly_id1, assembly_id2, limits, frhs, fwarm, c.key));
                        bil_contact.push_back(ci);
                    }
                }
            }
        }
    }

    /* signorini_coulomb_pyramid_model.rs:226-261.  The cache is rebuilt from
     * this step's contacts only: a ContactId (slotmap key) that disappears is
     * never issued again, so forgetting absent keys is equivalent. */
    void contact_cache_impulses() {
        contact_impulses_out.assign(contacts.size(), v3(0, 0, 0));
        if (contact_model == 1) {
            /* signorini_model.rs:285-297: the solved contacts are inserted, nothing is ever removed, so a
             * contact that was inactive this step keeps the impulse of the step it was last solved in;
             * contact_impulses_out reports the cache entry of every contact */
            for (size_t ci = 0; ci < contacts.size(); ++ci) {
                auto it = contacts[ci].key != 0 ? impulses.find(contacts[ci].key) : impulses.end();
                if (it != impulses.end()) contact_impulses_out[ci].x = it->second.x;
            }
        } else {
            /* The reference's cache (a SecondaryMap) never forgets.  A ContactId that leaves the narrow
             * phase is never issued again, so forgetting ABSENT ids is equivalent -- but a contact that is
             * still reported and merely not solved this step (its pair sleeps: it is filtered out of the
             * manifold list, mechanical_world.rs:287-300) must keep its entry, or the island would
             * cold-start when it wakes.  Those entries are carried over. */
            std::unordered_map<uint64_t, V3> old;
            old.swap(impulses);
            for (size_t ci = 0; ci < contacts.size(); ++ci) {
                const uint64_t key = contacts[ci].key;
                if (key == 0) continue;
                auto it = old.find(key);
                if (it != old.end()) impulses[key] = it->second; /* overwritten below if solved this step */
            }
            solved_scratch.assign(contacts.size(), 0);
            for (size_t ci : unig_contact) solved_scratch[ci] = 1;
            for (size_t ci : uni_contact) solved_scratch[ci] = 1;
            for (size_t ci = 0; ci < contacts.size(); ++ci)
                if (!solved_scratch[ci] && contacts[ci].key != 0) {
                    auto it = impulses.find(contacts[ci].key);
                    if (it != impulses.end()) contact_impulses_out[ci] = it->second;
                }
            /* entries of solved contacts are rebuilt from scratch below */
            for (size_t ci = 0; ci < contacts.size(); ++ci)
                if (solved_scratch[ci] && contacts[ci].key != 0) impulses.erase(contacts[ci].key);
        }
        for (size_t k = 0; k < contact_vel.unilateral_ground.size(); ++k) {
            const UnilateralGround& c = contact_vel.unilateral_ground[k];
            contact_impulses_out[unig_contact[k]].x = c.impulse;
            if (c.impulse_id != 0) impulses[c.impulse_id] = v3(c.impulse, 0, 0);
        }
        for (size_t k = 0; k < contact_vel.unilateral.size(); ++k) {
            const Unilateral& c = contact_vel.unilateral[k];
            contact_impulses_out[uni_contact[k]].x = c.impulse;
            if (c.impulse_id != 0) impulses[c.impulse_id] = v3(c.impulse, 0, 0);
        }
        size_t dim = 0, dim_all = 0;
        for (size_t k = 0; k < contact_vel.bilateral_ground.size(); ++k) {
            const BilateralGround& c = contact_vel.bilateral_ground[k];
            V3& o = contact_impulses_out[bilg_contact[k]];
            if (dim_all % 2 == 0) o.y = c.impulse; else o.z = c.impulse;
            ++dim_all;
            if (c.impulse_id != 0) {
                V3& e = impulses[c.impulse_id];
                if (dim % 2 == 0) e.y = c.impulse; else e.z = c.impulse;
                ++dim;
            }
        }
        for (size_t k = 0; k < contact_vel.bilateral.size(); ++k) {
            const Bilateral& c = contact_vel.bilateral[k];
            V3& o = contact_impulses_out[bil_contact[k]];
            if (dim_all % 2 == 0) o.y = c.impulse; else o.z = c.impulse;
            ++dim_all;
            if (c.impulse_id != 0) {
                V3& e = impulses[c.impulse_id];
                if (dim % 2 == 0) e.y = c.impulse; else e.z = c.impulse;
                ++dim;
            }
        }
    }

    /* ------------------------------------------------------------ SORProx */
    /* sor_prox.rs:345-435. */
    void warmstart_set(LinearConstraints& cs) {
        real* lam = mj_lambda_vel.data();
        const real* jac = jacobians.data();
        for (auto& c : cs.unilateral)
            if (c.impulse != (real)0) {
                axpyn(c.ndofs1, c.impulse, jac + c.wj_id1, lam + c.assembly_id1);
                axpyn(c.ndofs2, c.impulse, jac + c.wj_id2, lam + c.assembly_id2);
            }
        for (auto& c : cs.unilateral_ground)
            if (c.impulse != (real)0) axpyn(c.ndofs, c.impulse, jac + c.wj_id, lam + c.assembly_id);
        for (auto& c : cs.bilateral)
            if (c.impulse != (real)0) {
                axpyn(c.ndofs1, c.impulse, jac + c.wj_id1, lam + c.assembly_id1);
                axpyn(c.ndofs2, c.impulse, jac + c.wj_id2, lam + c.assembly_id2);
            }
        for (auto& c : cs.bilateral_ground)
            if (c.impulse != (real)0) axpyn(c.ndofs, c.impulse, jac + c.wj_id, lam + c.assembly_id);
    }
    static real clampv(real v, real lo, real hi) { /* na::clamp */
        return v > lo ? (v < hi ? v : hi) : lo;
    }
    /* sor_prox.rs:112-157, 232-343. */
    void step_bilateral(LinearConstraints& cs) {
        real* lam = mj_lambda_vel.data();
        const real* jac = jacobians.data();
        for (auto& c : cs.bilateral) {
            real min_impulse, max_impulse;
            if (c.limits.kind == LIMIT_INDEPENDENT) {
                min_impulse = c.limits.min;
                max_impulse = c.limits.max;
            } else {
                real impulse = cs.unilateral[c.limits.dependency].impulse;
                if (impulse == (real)0) {
                    if (c.impulse != (real)0) {
                        axpyn(c.ndofs1, -c.impulse, jac + c.wj_id1, lam + c.assembly_id1);
                        axpyn(c.ndofs2, -c.impulse, jac + c.wj_id2, lam + c.assembly_id2);
                        c.impulse = 0;
                    }
                    continue;
                }
                max_impulse = c.limits.coeff * impulse;
                min_impulse = -max_impulse;
            }
            real dimpulse = dotn(c.ndofs1, jac + c.j_id1, lam + c.assembly_id1) + dotn(c.ndofs2, jac + c.j_id2, lam + c.assembly_id2) + c.rhs;
            real new_impulse = clampv(c.impulse - c.r * dimpulse, min_impulse, max_impulse);
            real dlambda = new_impulse - c.impulse;
            c.impulse = new_impulse;
            axpyn(c.ndofs1, dlambda, jac + c.wj_id1, lam + c.assembly_id1);
            axpyn(c.ndofs2, dlambda, jac + c.wj_id2, lam + c.assembly_id2);
        }
        for (auto& c : cs.bilateral_ground) {
            real min_impulse, max_impulse;
            if (c.limits.kind == LIMIT_INDEPENDENT) {
                min_impulse = c.limits.min;
                max_impulse = c.limits.max;
            } else {
                real impulse = cs.unilateral_ground[c.limits.dependency].impulse;
                if (impulse == (real)0) {
                    if (c.impulse != (real)0) {
                        axpyn(c.ndofs, -c.impulse, jac + c.wj_id, lam + c.assembly_id);
                        c.impulse = 0;
                    }
                    continue;
                }
                max_impulse = c.limits.coeff * impulse;
                min_impulse = -max_impulse;
            }
            real dimpulse = dotn(c.ndofs, jac + c.j_id, lam + c.assembly_id) + c.rhs;
            real new_impulse = clampv(c.impulse - c.r * dimpulse, min_impulse, max_impulse);
            real dlambda = new_impulse - c.impulse;
            c.impulse = new_impulse;
            axpyn(c.ndofs, dlambda, jac + c.wj_id, lam + c.assembly_id);
        }
    }
    /* sor_prox.rs:82-110, 181-230. */
    void step_unilateral(LinearConstraints& cs) {
        real* lam = mj_lambda_vel.data();
        const real* jac = jacobians.data();
        for (auto& c : cs.unilateral) {
            real dimpulse = dotn(c.ndofs1, jac + c.j_id1, lam + c.assembly_id1) + dotn(c.ndofs2, jac + c.j_id2, lam + c.assembly_id2) + c.rhs;
            real new_impulse = std::max(c.impulse - c.r * dimpulse, (real)0);
            real dlambda = new_impulse - c.impulse;
            c.impulse = new_impulse;
            axpyn(c.ndofs1, dlambda, jac + c.wj_id1, lam + c.assembly_id1);
            axpyn(c.ndofs2, dlambda, jac + c.wj_id2, lam + c.assembly_id2);
        }
        for (auto& c : cs.unilateral_ground) {
            real dimpulse = dotn(c.ndofs, jac + c.j_id, lam + c.assembly_id) + c.rhs;
            real new_impulse = std::max(c.impulse - c.r * dimpulse, (real)0);
            real dlambda = new_impulse - c.impulse;
            c.impulse = new_impulse;
            axpyn(c.ndofs, dlambda, jac + c.wj_id, lam + c.assembly_id);
        }
    }
    /* SORProx::solve, sor_prox.rs:48-80 and step :159-179. */
    void sor_prox_solve(size_t max_iter) {
        warmstart_set(contact_vel);
        warmstart_set(joint_vel);
        for (Multibody& mb : mbs) /* sor_prox.rs:60-65 */
            if (mb.has_active_internal_constraints()) mb.warmstart_internal_velocity_constraints(&mj_lambda_vel[mb.companion_id]);
        for (size_t it = 0; it < max_iter; ++it) {
            step_bilateral(joint_vel);
            step_bilateral(contact_vel);
            for (Multibody& mb : mbs) /* sor_prox.rs:170-175 */
                if (mb.has_active_internal_constraints()) mb.step_solve_internal_velocity_constraints(&mj_lambda_vel[mb.companion_id]);
            step_unilateral(joint_vel);
            step_unilateral(contact_vel);
        }
    }

    /* ----------------------------------------------------- NonlinearSORProx */
    std::vector<real> pos_jacobians; /* the cloned jacobian buffer (moreau_jean_solver.rs:292) */

    real clamp_rhs(real rhs, bool is_angular) const { /* nonlinear_sor_prox.rs:296-309 */
        if (is_angular)
            return std::max((rhs + (real)params.allowed_angular_error) * (real)params.erp,
                            -(real)params.max_angular_correction);
        return std::max((rhs + (real)params.allowed_linear_error) * (real)params.erp,
                        -(real)params.max_linear_correction);
    }
    /* nonlinear_sor_prox.rs:78-119. */
    void solve_generic(GenericNonlinear& c) {
        real rhs = clamp_rhs(c.rhs, c.is_angular);
        if (rhs < (real)0) {
            real impulse = -rhs * c.r;
            for (size_t k = 0; k < c.dim1; ++k) pos_jacobians[c.wj_id1 + k] *= impulse;
            for (size_t k = 0; k < c.dim2; ++k) pos_jacobians[c.wj_id2 + k] *= impulse;
            if (c.dim1 != 0) apply_body_displacement(c.body1, &pos_jacobians[c.wj_id1]);
            if (c.dim2 != 0) apply_body_displacement(c.body2, &pos_jacobians[c.wj_id2]);
        }
    }
    /* nonlinear_sor_prox.rs:156-294. */
    bool update_contact_constraint(NonlinearUnilateral& c) {
        const Body& body1 = bodies[c.body1];
        const Body& body2 = bodies[c.body2];
        Iso pos1 = body1.position * c.coll1_wrt_body;
        Iso pos2 = body2.position * c.coll2_wrt_body;
        Contact contact;
        if (!kinematic_contact(c, pos1, pos2, &contact)) return false;
        c.rhs = clamp_rhs(-contact.depth, false);
        if (c.rhs >= (real)0) return false;
        real inv_r = 0;
        size_t j_id1 = c.ndofs1 + c.ndofs2;
        size_t j_id2 = c.ndofs1 * 2 + c.ndofs2;
        if (pos_jacobians.size() < j_id2 + 2 * c.ndofs2 + 6) pos_jacobians.resize(j_id2 + 2 * c.ndofs2 + 6, 0);
        if (c.ndofs1 != 0)
            fill_constraint_geometry(body1, contact.world1, fd_linear(-contact.normal), j_id1, 0, pos_jacobians.data(),
                                     &inv_r, nullptr, nullptr);
        if (c.ndofs2 != 0)
            fill_constraint_geometry(body2, contact.world2, fd_linear(contact.normal), j_id2, c.ndofs1,
                                     pos_jacobians.data(), &inv_r, nullptr, nullptr);
        if (inv_r == (real)0) return false;
        c.r = (real)1 / inv_r;
        return true;
    }
    /* nonlinear_sor_prox.rs:121-154. */
    void solve_unilateral_position(NonlinearUnilateral& c) {
        if (update_contact_constraint(c)) {
            real impulse = -c.rhs * c.r;
            for (size_t k = 0; k < c.ndofs1; ++k) pos_jacobians[k] *= impulse;
            for (size_t k = 0; k < c.ndofs2; ++k) pos_jacobians[c.ndofs1 + k] *= impulse;
            if (c.ndofs1 != 0) apply_body_displacement(c.body1, &pos_jacobians[0]);
            if (c.ndofs2 != 0) apply_body_displacement(c.body2, &pos_jacobians[c.ndofs1]);
        }
    }
    /* NonlinearSORProx::solve, nonlinear_sor_prox.rs:17-55. */
    void nonlinear_sor_prox_solve(size_t max_iter) {
        pos_jacobians = jacobians; /* moreau_jean_solver.rs:292 */
        if (pos_jacobians.size() < 64) pos_jacobians.resize(64, 0);
        for (size_t it = 0; it < max_iter; ++it) {
            for (size_t ji : active_joints) {
                Joint& j = joints[ji];
                size_t n = joint_num_position_constraints(j);
                for (size_t i = 0; i < n; ++i) {
                    GenericNonlinear g;
                    if (joint_position_constraint(j, i, &g)) solve_generic(g);
                }
            }
            for (Multibody& mb : mbs) /* nonlinear_sor_prox.rs:40-44 */
                if (mb.has_active_internal_constraints()) mb.step_solve_internal_position_constraints(params, bodies);
            for (NonlinearUnilateral& c : contact_pos) solve_unilateral_position(c);
        }
    }

    /* ------------------------------------------------- MoreauJeanSolver */
    /* assemble_system, moreau_jean_solver.rs:129-260. */
    void assemble_system() {
        size_t system_ndofs = 0;
        for (int h : island) {
            Body& b = bodies[h];
            b.companion_id = system_ndofs;
            system_ndofs += 6;
        }
        for (Multibody& mb : mbs) { /* a multibody is one body of the island with ndofs dofs; its links share its id */
            mb.companion_id = system_ndofs;
            system_ndofs += mb.ndofs;
            for (const MbLink& l : mb.rbs)
                if (l.body >= 0) bodies[l.body].companion_id = mb.companion_id;
        }
        mj_lambda_vel.assign(system_ndofs, 0); /* resize_buffers :322-326 */
        ext_vels.assign(system_ndofs, 0);
        contact_vel.clear();
        joint_vel.clear();
        contact_pos.clear();
        const real dt = params.dt;
        for (int h : island) { /* :166-174: ext_vels = dt * acc + 0 * ext_vels */
            const Body& b = bodies[h];
            real* e = &ext_vels[b.companion_id];
            const S6& a = b.acceleration;
            const real acc[6] = {a.lin.x, a.lin.y, a.lin.z, a.ang.x, a.ang.y, a.ang.z};
            for (int k = 0; k < 6; ++k) e[k] = dt * acc[k];
        }
        for (const Multibody& mb : mbs)
            for (size_t k = 0; k < mb.ndofs; ++k) ext_vels[mb.companion_id + k] = dt * mb.accelerations[k];
        /* jacobian sizes :181-216 */
        size_t jacobian_sz = 0, ground_jacobian_sz = 0;
        static const size_t max_rows[NB2_JOINT_TYPE_COUNT] = {3, 5, 7, 4, 3, 4, 4, 4, 6, 3};
        for (size_t ji : active_joints) {
            const Joint& j = joints[ji];
            size_t nd1 = bodies[j.rec.body1].status_dependent_ndofs(), nd2 = bodies[j.rec.body2].status_dependent_ndofs();
            size_t sz = max_rows[j.rec.type] * 2 * (nd1 + nd2);
            if (nd1 == 0 || nd2 == 0) ground_jacobian_sz += sz; else jacobian_sz += sz;
        }
        for (const nb2_manifold& m : manifolds) {
            size_t nd1 = bodies[m.body1].status_dependent_ndofs(), nd2 = bodies[m.body2].status_dependent_ndofs();
            size_t sz = 3 * m.num_contacts * (nd1 + nd2) * 2;
            if (nd1 == 0 || nd2 == 0) ground_jacobian_sz += sz; else jacobian_sz += sz;
        }
        jacobians.assign(jacobian_sz + ground_jacobian_sz, 0);
        size_t j_id = 0, ground_j_id = jacobian_sz;
        for (size_t ji : active_joints) joint_velocity_constraints(joints[ji], &ground_j_id, &j_id);
        contact_constraints(&ground_j_id, &j_id);
        for (Multibody& mb : mbs) /* :254-259 */
            if (mb.has_active_internal_constraints()) mb.setup_internal_velocity_constraints(&ext_vels[mb.companion_id], params);
    }
    /* update_velocities_and_integrate, moreau_jean_solver.rs:328-347. */
    void update_velocities_and_integrate() {
        for (int h : island) {
            Body& b = bodies[h];
            const real* e = &ext_vels[b.companion_id];
            const real* l = &mj_lambda_vel[b.companion_id];
            real v[6] = {b.velocity.lin.x, b.velocity.lin.y, b.velocity.lin.z,
                         b.velocity.ang.x, b.velocity.ang.y, b.velocity.ang.z};
            for (int k = 0; k < 6; ++k) v[k] += e[k];
            for (int k = 0; k < 6; ++k) v[k] += l[k];
            b.velocity.lin = v3(v[0], v[1], v[2]);
            b.velocity.ang = v3(v[3], v[4], v[5]);
            b.integrate(params.dt);
        }
        for (Multibody& mb : mbs) {
            for (size_t k = 0; k < mb.ndofs; ++k) mb.velocities[k] += ext_vels[mb.companion_id + k];
            for (size_t k = 0; k < mb.ndofs; ++k) mb.velocities[k] += mj_lambda_vel[mb.companion_id + k];
            mb.integrate(params.dt);
        }
    }

    double t_assembly = 0, t_velocity = 0, t_update = 0, t_position = 0, t_step = 0;

    /* MechanicalWorld::step sequencing around the solver (mechanical_world.rs:230-243,
     * 274-346) + MoreauJeanSolver::step (moreau_jean_solver.rs:47-90). */
    int step() {
        using clk = std::chrono::steady_clock;
        auto T0 = clk::now();
        const real dt = params.dt;
        /* :230-243.  update_dynamics at :233 is a no-op unless dirty; the dirty
         * state was consumed at the end of the previous step (:343-346) on the
         * same pose/velocity, so recomputing here is equivalent. */
        for (Body& b : bodies) b.update_dynamics(dt);
        V3 g = v3(params.gravity[0], params.gravity[1], params.gravity[2]);
        for (Body& b : bodies) b.update_acceleration(g);
        for (Multibody& mb : mbs) { /* mechanical_world.rs:230-243 */
            mb.update_kinematics(bodies);
            mb.update_dynamics(dt, bodies);
            mb.update_acceleration(g);
        }
        /* :264-279: active_bodies = the non-kinematic members of the islands that stay awake; the caller
         * ran update_activation (or sleeping is off and every dynamic body is active) */
        island.clear();
        for (size_t i = 0; i < bodies.size(); ++i)
            if (bodies[i].status == NB2_BODY_DYNAMIC && bodies[i].is_active()) island.push_back((int)i);
        /* :287-300 contact manifolds of the step */
        manifolds.clear();
        for (const nb2_manifold& m : uploaded_manifolds) {
            const Body& b1 = bodies[m.body1];
            const Body& b2 = bodies[m.body2];
            if (m.num_contacts > 0 && b1.status != NB2_BODY_DISABLED && b2.status != NB2_BODY_DISABLED &&
                ((b1.status_dependent_ndofs() != 0 && b1.is_active()) || (b2.status_dependent_ndofs() != 0 && b2.is_active())))
                manifolds.push_back(m);
        }
        active_joints.clear();
        for (size_t i = 0; i < joints.size(); ++i)
            if (!joints[i].rec.broken && joint_is_active(joints[i])) active_joints.push_back(i);
        for (Body& b : bodies) b.companion_id = 0; /* :307-313 */

        auto T1 = clk::now();
        assemble_system();
        auto T2 = clk::now();
        sor_prox_solve(params.max_velocity_iterations);
        contact_cache_impulses(); /* cache_impulses :306-320 */
        for (size_t ji : active_joints) joint_cache_impulses(joints[ji]);
        auto T3 = clk::now();
        update_velocities_and_integrate();
        auto T4 = clk::now();
        nonlinear_sor_prox_solve(params.max_position_iterations);
        auto T5 = clk::now();
        for (Body& b : bodies) /* :328-332 */
            if (b.status == NB2_BODY_KINEMATIC) b.integrate(dt);
        for (Multibody& mb : mbs) { /* mechanical_world.rs:343-346: kinematics and dynamics after the resolution */
            mb.update_kinematics(bodies);
            mb.update_dynamics(dt, bodies);
        }
        auto ms = [](clk::time_point a, clk::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        t_assembly = ms(T1, T2);
        t_velocity = ms(T2, T3);
        t_update = ms(T3, T4);
        t_position = ms(T4, T5);
        t_step = ms(T0, clk::now());
        compute_residual(); /* diagnostics only: outside every timed stage */
        return NB2_OK;
    }

    /* MoreauJeanSolver::step_ccd (moreau_jean_solver.rs:94-127) with the island / manifold / joint
     * selection of the regular step: assemble, position constraints first, velocity constraints, then
     * update_velocities_and_integrate.  No cache_impulses, no kinematic integration. */
    int step_ccd() {
        const real dt = params.dt;
        for (Body& b : bodies) b.update_dynamics(dt);
        V3 g = v3(params.gravity[0], params.gravity[1], params.gravity[2]);
        for (Body& b : bodies) b.update_acceleration(g);
        for (Multibody& mb : mbs) { /* mechanical_world.rs:230-243 */
            mb.update_kinematics(bodies);
            mb.update_dynamics(dt, bodies);
            mb.update_acceleration(g);
        }
        island.clear();
        for (size_t i = 0; i < bodies.size(); ++i)
            if (bodies[i].status == NB2_BODY_DYNAMIC && bodies[i].is_active()) island.push_back((int)i);
        manifolds.clear();
        for (const nb2_manifold& m : uploaded_manifolds) {
            const Body& b1 = bodies[m.body1];
            const Body& b2 = bodies[m.body2];
            if (m.num_contacts > 0 && b1.status != NB2_BODY_DISABLED && b2.status != NB2_BODY_DISABLED &&
                ((b1.status_dependent_ndofs() != 0 && b1.is_active()) || (b2.status_dependent_ndofs() != 0 && b2.is_active())))
                manifolds.push_back(m);
        }
        active_joints.clear();
        for (size_t i = 0; i < joints.size(); ++i)
            if (!joints[i].rec.broken && joint_is_active(joints[i])) active_joints.push_back(i);
        for (Body& b : bodies) b.companion_id = 0;
        assemble_system();
        nonlinear_sor_prox_solve(params.max_position_iterations); /* :121 solve_position_constraints */
        sor_prox_solve(params.max_velocity_iterations);           /* :126 solve_velocity_constraints */
        update_velocities_and_integrate();                         /* :127 */
        for (Multibody& mb : mbs) {
            mb.update_kinematics(bodies);
            mb.update_dynamics(dt, bodies);
        }
        compute_residual();
        return NB2_OK;
    }

    /* Diagnostics shared with the CUDA path's nb2_get_stats (not part of the
     * reference): natural-map residual of every velocity row at the final
     * iterate, evaluated Jacobi-style. */
    real res_max = 0;
    double res_sq = 0;
    size_t res_n = 0;
    void compute_residual() {
        res_max = 0;
        res_sq = 0;
        res_n = 0;
        const real* lam = mj_lambda_vel.data();
        const real* jac = jacobians.data();
        auto acc = [&](real d) {
            d = std::fabs(d);
            if (d > res_max) res_max = d;
            res_sq += (double)d * d;
            ++res_n;
        };
        for (LinearConstraints* cs : {&joint_vel, &contact_vel}) {
            for (auto& c : cs->unilateral) {
                real w = dotn(c.ndofs1, jac + c.j_id1, lam + c.assembly_id1) + dotn(c.ndofs2, jac + c.j_id2, lam + c.assembly_id2) + c.rhs;
                acc(std::max(c.impulse - c.r * w, (real)0) - c.impulse);
            }
            for (auto& c : cs->unilateral_ground) {
                real w = dotn(c.ndofs, jac + c.j_id, lam + c.assembly_id) + c.rhs;
                acc(std::max(c.impulse - c.r * w, (real)0) - c.impulse);
            }
            for (auto& c : cs->bilateral) {
                real lo = c.limits.min, hi = c.limits.max;
                if (c.limits.kind == LIMIT_DEPENDENT) {
                    hi = c.limits.coeff * cs->unilateral[c.limits.dependency].impulse;
                    lo = -hi;
                }
                real w = dotn(c.ndofs1, jac + c.j_id1, lam + c.assembly_id1) + dotn(c.ndofs2, jac + c.j_id2, lam + c.assembly_id2) + c.rhs;
                acc(clampv(c.impulse - c.r * w, lo, hi) - c.impulse);
            }
            for (auto& c : cs->bilateral_ground) {
                real lo = c.limits.min, hi = c.limits.max;
                if (c.limits.kind == LIMIT_DEPENDENT) {
                    hi = c.limits.coeff * cs->unilateral_ground[c.limits.dependency].impulse;
                    lo = -hi;
                }
                real w = dotn(c.ndofs, jac + c.j_id, lam + c.assembly_id) + c.rhs;
                acc(clampv(c.impulse - c.r * w, lo, hi) - c.impulse);
            }
        }
    }
    void fill_stats(nb2_stats* s) {
        std::memset(s, 0, sizeof(*s));
        s->n_bodies = (uint32_t)bodies.size();
        s->n_dynamic_bodies = (uint32_t)island.size();
        s->n_manifolds = (uint32_t)manifolds.size();
        s->n_contacts = (uint32_t)contacts.size();
        s->n_joints = (uint32_t)joints.size();
        s->n_rows_two_body = (uint32_t)(contact_vel.unilateral.size() + contact_vel.bilateral.size() +
                                        joint_vel.unilateral.size() + joint_vel.bilateral.size());
        s->n_rows_ground = (uint32_t)(contact_vel.unilateral_ground.size() + contact_vel.bilateral_ground.size() +
                                      joint_vel.unilateral_ground.size() + joint_vel.bilateral_ground.size());
        for (const Joint& j : joints) s->n_broken_joints += j.rec.broken ? 1 : 0;
        s->residual_max = res_max;
        s->residual_rms = res_n ? (float)std::sqrt(res_sq / (double)res_n) : 0.f;
        /* max penetration at the final poses */
        real pen = -REAL_MAX;
        for (NonlinearUnilateral& c : contact_pos) {
            Contact ct;
            Iso p1 = bodies[c.body1].position * c.coll1_wrt_body;
            Iso p2 = bodies[c.body2].position * c.coll2_wrt_body;
            if (kinematic_contact(c, p1, p2, &ct)) pen = std::max(pen, ct.depth);
        }
        s->max_penetration = contact_pos.empty() ? 0.f : (float)pen;
        double ke = 0;
        for (const Body& b : bodies) {
            if (b.status != NB2_BODY_DYNAMIC) continue;
            M3 rot = to_rotation_matrix(b.position.r);
            M3 iw = (rot * b.local_inertia.angular) * transpose(rot);
            V3 l = iw * b.velocity.ang;
            ke += 0.5 * (double)b.local_inertia.linear * (double)norm_squared(b.velocity.lin) +
                  0.5 * (double)dot(b.velocity.ang, l);
            const real chk[13] = {b.position.t.x, b.position.t.y, b.position.t.z, b.position.r.i, b.position.r.j,
                                  b.position.r.k, b.position.r.w, b.velocity.lin.x, b.velocity.lin.y, b.velocity.lin.z,
                                  b.velocity.ang.x, b.velocity.ang.y, b.velocity.ang.z};
            for (real x : chk)
                if (!std::isfinite(x)) {
                    s->non_finite++;
                    break;
                }
        }
        s->kinetic_energy = (float)ke;
        s->t_assembly_ms = (float)t_assembly;
        s->t_velocity_resolution_ms = (float)t_velocity;
        s->t_velocity_update_ms = (float)t_update;
        s->t_position_resolution_ms = (float)t_position;
        s->t_step_ms = (float)t_step;
    }
};

}  // namespace

/* ------------------------------------------------------------------- C API */
extern "C" {

void* nbo_create(void) { return new World(); }
void nbo_destroy(void* w) { delete (World*)w; }

int nbo_set_params(void* wp, const nb2_params* p) {
    World* w = (World*)wp;
    if (!p || p->dt < 0) return NB2_ERR_INVALID_ARGUMENT;
    w->params = *p;
    w->inv_dt = p->dt == 0.f ? (real)0 : (real)1 / (real)p->dt; /* integration_parameters.rs:141-153 */
    return NB2_OK;
}

int nbo_upload_bodies(void* wp, const nb2_body* in, uint32_t n) {
    World* w = (World*)wp;
    w->bodies.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        const nb2_body& s = in[i];
        Body& b = w->bodies[i];
        b.position = Iso{ld3(s.position), ldq(s.position + 3)};
        b.velocity = S6{ld3(s.velocity), ld3(s.velocity + 3)};
        b.local_com = ld3(s.local_com);
        b.local_inertia.linear = s.mass;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) b.local_inertia.angular.m[r][c] = s.local_inertia[r * 3 + c];
        b.external_forces = S6{ld3(s.external_forces), ld3(s.external_forces + 3)};
        b.linear_damping = s.linear_damping;
        b.angular_damping = s.angular_damping;
        b.max_linear_velocity = s.max_linear_velocity;
        b.max_angular_velocity = s.max_angular_velocity;
        for (int k = 0; k < 6; ++k) b.jacobian_mask[k] = s.jacobian_mask[k];
        b.status = s.status;
        b.gravity_enabled = (s.flags & NB2_BODY_FLAG_GRAVITY) != 0;
        b.com = transform_point(b.position, b.local_com);
        b.inertia = Inertia{0, m3_zero()};
        b.augmented_mass = b.inertia;
        b.inv_augmented_mass = b.inertia;
        b.acceleration = S6{v3(0, 0, 0), v3(0, 0, 0)};
        b.companion_id = 0;
        b.mb = nullptr;
        b.mb_link = -1;
        b.handle = (int)i;
    }
    w->mbs.clear();
    w->mb_link_recs.clear();
    return NB2_OK;
}

/* MultibodyDesc::build (multibody.rs:1433-1470): links in the uploaded order, add_link :180-254. */
int nbo_upload_multibodies(void* wp, const nb2_multibody* mb_in, uint32_t n_mb, const nb2_mb_link* links, uint32_t n_links) {
    World* w = (World*)wp;
    w->mbs.assign(n_mb, Multibody());
    w->mb_link_recs.assign(links, links + n_links);
    for (uint32_t m = 0; m < n_mb; ++m) {
        Multibody& mb = w->mbs[m];
        const nb2_multibody& rec = mb_in[m];
        if ((size_t)rec.first_link + rec.n_links > n_links || rec.n_links == 0) return NB2_ERR_BAD_INDEX;
        mb.gravity_enabled = (rec.flags & NB2_BODY_FLAG_GRAVITY) != 0;
        for (uint32_t k = 0; k < rec.n_links; ++k) {
            const nb2_mb_link& s = links[rec.first_link + k];
            if (s.joint_type >= NB2_MBJ_TYPE_COUNT || s.parent >= (int)k || (k == 0) != (s.parent < 0)) return NB2_ERR_INVALID_ARGUMENT;
            if (s.body < 0 || (size_t)s.body >= w->bodies.size() || w->bodies[s.body].status != NB2_BODY_MULTIBODY_LINK)
                return NB2_ERR_BAD_INDEX;
            MbLink l;
            l.parent = s.parent;
            l.type = s.joint_type;
            l.flags = s.flags;
            l.body = s.body;
            l.parent_shift = ld3(s.parent_shift);
            l.body_shift = ld3(s.body_shift);
            l.axis = ld3(s.axis);
            const Body& pb = w->bodies[s.body];
            l.local_com = pb.local_com;
            l.local_inertia = pb.local_inertia;
            l.min_pos = s.min_pos;
            l.max_pos = s.max_pos;
            l.motor_velocity = s.motor_velocity;
            l.motor_max_velocity = s.motor_max_velocity;
            l.motor_max_force = s.motor_max_force;
            l.free_pos = Iso{v3(0, 0, 0), quat_identity()};
            l.rot = quat_identity();
            switch (s.joint_type) {
                case NB2_MBJ_FREE:
                case NB2_MBJ_FIXED: l.free_pos = Iso{ld3(s.coords), ldq(s.coords + 3)}; break;
                case NB2_MBJ_BALL: l.rot = ldq(s.coords); break;
                case NB2_MBJ_REVOLUTE:
                    l.coord = s.coords[0];
                    l.rot = from_axis_angle(l.axis, l.coord);
                    break;
                default: l.coord = s.coords[0]; break;
            }
            l.ndofs = MbLink::ndofs_of(s.joint_type);
            l.assembly_id = mb.velocities.size();
            l.impulse_id = mb.impulses.size();
            for (size_t d = 0; d < l.ndofs; ++d) {
                mb.velocities.push_back(s.velocity[d]);
                mb.damping.push_back(s.damping[d]);
                mb.accelerations.push_back(0);
                mb.forces.push_back(0);
            }
            for (size_t d = 0; d < l.ndofs * 3; ++d) mb.impulses.push_back(l.ndofs == 1 ? s.impulses[d] : 0); /* Joint::nimpulses */
            mb.ndofs += l.ndofs;
            mb.rbs.push_back(l);
        }
        if (mb.ndofs > NB2_MB_MAX_DOFS) return NB2_ERR_UNSUPPORTED;
    }
    for (uint32_t m = 0; m < n_mb; ++m) {
        Multibody& mb = w->mbs[m];
        for (size_t k = 0; k < mb.rbs.size(); ++k) {
            Body& pb = w->bodies[mb.rbs[k].body];
            pb.mb = &mb;
            pb.mb_link = (int)k;
            pb.handle = (int)(w->bodies.size() + m);
        }
        mb.update_kinematics(w->bodies);
        mb.update_dynamics(w->params.dt, w->bodies);
    }
    return NB2_OK;
}
int nbo_download_multibody_links(void* wp, nb2_mb_link* out, uint32_t n) {
    World* w = (World*)wp;
    if (n > w->mb_link_recs.size()) return NB2_ERR_BAD_INDEX;
    for (uint32_t i = 0; i < n; ++i) {
        nb2_mb_link r = w->mb_link_recs[i];
        const Multibody& mb = w->mbs[r.multibody];
        uint32_t first = 0; /* link index within its multibody */
        for (uint32_t k = 0; k < i; ++k)
            if (w->mb_link_recs[k].multibody == r.multibody) ++first;
        const MbLink& l = mb.rbs[first];
        switch (l.type) {
            case NB2_MBJ_FREE:
            case NB2_MBJ_FIXED:
                r.coords[0] = l.free_pos.t.x; r.coords[1] = l.free_pos.t.y; r.coords[2] = l.free_pos.t.z;
                r.coords[3] = l.free_pos.r.i; r.coords[4] = l.free_pos.r.j; r.coords[5] = l.free_pos.r.k; r.coords[6] = l.free_pos.r.w;
                break;
            case NB2_MBJ_BALL:
                r.coords[0] = l.rot.i; r.coords[1] = l.rot.j; r.coords[2] = l.rot.k; r.coords[3] = l.rot.w;
                break;
            default: r.coords[0] = l.coord; break;
        }
        for (size_t d = 0; d < l.ndofs; ++d) r.velocity[d] = mb.velocities[l.assembly_id + d];
        if (l.ndofs == 1) {
            /* the rows' impulses reach Multibody::impulses at the next setup (multibody.rs:1046-1053): report the
             * latest ones */
            std::vector<real> imp(mb.impulses.begin() + l.impulse_id, mb.impulses.begin() + l.impulse_id + 3);
            for (const UnilateralGround& c : mb.internal.unilateral_ground)
                if (c.impulse_id >= l.impulse_id && c.impulse_id < l.impulse_id + 3) imp[c.impulse_id - l.impulse_id] = c.impulse;
            for (const BilateralGround& c : mb.internal.bilateral_ground)
                if (c.impulse_id >= l.impulse_id && c.impulse_id < l.impulse_id + 3) imp[c.impulse_id - l.impulse_id] = c.impulse;
            for (int d = 0; d < 3; ++d) r.impulses[d] = imp[d];
        }
        out[i] = r;
    }
    return NB2_OK;
}

int nbo_upload_body_states(void* wp, const nb2_body_state* in, uint32_t first, uint32_t n) {
    World* w = (World*)wp;
    if ((size_t)first + n > w->bodies.size()) return NB2_ERR_BAD_INDEX;
    for (uint32_t i = 0; i < n; ++i) {
        Body& b = w->bodies[first + i];
        b.set_position(Iso{ld3(in[i].position), ldq(in[i].position + 3)});
        b.velocity = S6{ld3(in[i].velocity), ld3(in[i].velocity + 3)};
    }
    return NB2_OK;
}

int nbo_upload_manifolds(void* wp, const nb2_manifold* m, uint32_t nm, const nb2_contact* c, uint32_t nc) {
    World* w = (World*)wp;
    for (uint32_t i = 0; i < nm; ++i) {
        if (m[i].body1 < 0 || m[i].body2 < 0 || (size_t)m[i].body1 >= w->bodies.size() ||
            (size_t)m[i].body2 >= w->bodies.size() || (uint64_t)m[i].first_contact + m[i].num_contacts > nc)
            return NB2_ERR_BAD_INDEX;
    }
    w->uploaded_manifolds.assign(m, m + nm);
    w->manifolds.assign(m, m + nm);
    w->contacts.assign(c, c + nc);
    return NB2_OK;
}

/* Sleeping: nb2_activation records (threshold < 0 = None, energy 0 = asleep). */
int nbo_upload_activation(void* wp, const nb2_activation* a, uint32_t n) {
    World* w = (World*)wp;
    if (n != w->bodies.size()) return NB2_ERR_INVALID_ARGUMENT;
    for (uint32_t i = 0; i < n; ++i) {
        w->bodies[i].act_threshold = a[i].threshold;
        w->bodies[i].act_energy = a[i].energy;
    }
    w->sleeping_enabled = true;
    return NB2_OK;
}
int nbo_update_activation(void* wp, float mix_factor, const int32_t* to_activate, uint32_t n) {
    World* w = (World*)wp;
    if (!w->sleeping_enabled) return NB2_ERR_NOT_READY;
    w->update_activation((real)mix_factor, to_activate, n);
    return NB2_OK;
}
int nbo_download_activation(void* wp, nb2_activation* out, uint32_t n) {
    World* w = (World*)wp;
    if (n > w->bodies.size()) return NB2_ERR_BAD_INDEX;
    for (uint32_t i = 0; i < n; ++i) {
        out[i].threshold = (float)w->bodies[i].act_threshold;
        out[i].energy = (float)w->bodies[i].act_energy;
    }
    return NB2_OK;
}

int nbo_upload_joints(void* wp, const nb2_joint* j, uint32_t n) {
    World* w = (World*)wp;
    w->joints.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        if (j[i].body1 < 0 || j[i].body2 < 0 || (size_t)j[i].body1 >= w->bodies.size() ||
            (size_t)j[i].body2 >= w->bodies.size() || j[i].type >= NB2_JOINT_TYPE_COUNT)
            return NB2_ERR_BAD_INDEX;
        w->joints[i] = Joint();
        w->joints[i].rec = j[i];
    }
    return NB2_OK;
}

int nbo_set_contact_model(void* wp, int model) {
    if (!wp || (model != 0 && model != 1)) return -1;
    World* w = (World*)wp;
    if (w->contact_model != model) w->impulses.clear(); /* a fresh ContactModel starts with an empty cache */
    w->contact_model = model;
    return 0;
}
int nbo_clear_impulse_cache(void* wp) {
    ((World*)wp)->impulses.clear();
    return NB2_OK;
}

int nbo_step(void* wp) { return ((World*)wp)->step(); }
int nbo_step_ccd(void* wp) { return ((World*)wp)->step_ccd(); }

int nbo_download_body_states(void* wp, nb2_body_state* out, uint32_t first, uint32_t n) {
    World* w = (World*)wp;
    if ((size_t)first + n > w->bodies.size()) return NB2_ERR_BAD_INDEX;
    for (uint32_t i = 0; i < n; ++i) {
        const Body& b = w->bodies[first + i];
        float* p = out[i].position;
        p[0] = b.position.t.x; p[1] = b.position.t.y; p[2] = b.position.t.z;
        p[3] = b.position.r.i; p[4] = b.position.r.j; p[5] = b.position.r.k; p[6] = b.position.r.w;
        float* v = out[i].velocity;
        v[0] = b.velocity.lin.x; v[1] = b.velocity.lin.y; v[2] = b.velocity.lin.z;
        v[3] = b.velocity.ang.x; v[4] = b.velocity.ang.y; v[5] = b.velocity.ang.z;
    }
    return NB2_OK;
}

int nbo_download_contact_impulses(void* wp, float* out3, uint32_t n) {
    World* w = (World*)wp;
    if (n > w->contact_impulses_out.size()) return NB2_ERR_BAD_INDEX;
    for (uint32_t i = 0; i < n; ++i) {
        out3[3 * i + 0] = w->contact_impulses_out[i].x;
        out3[3 * i + 1] = w->contact_impulses_out[i].y;
        out3[3 * i + 2] = w->contact_impulses_out[i].z;
    }
    return NB2_OK;
}

int nbo_download_joints(void* wp, nb2_joint* out, uint32_t n) {
    World* w = (World*)wp;
    if (n > w->joints.size()) return NB2_ERR_BAD_INDEX;
    for (uint32_t i = 0; i < n; ++i) out[i] = w->joints[i].rec;
    return NB2_OK;
}

int nbo_get_stats(void* wp, nb2_stats* out) {
    ((World*)wp)->fill_stats(out);
    return NB2_OK;
}

/* Debug taps used by the unit tests of the restatement itself. */
int nbo_debug_body_dynamics(void* wp, uint32_t i, float* inv_aug10, float* acc6, float* com3) {
    World* w = (World*)wp;
    if (i >= w->bodies.size()) return NB2_ERR_BAD_INDEX;
    const Body& b = w->bodies[i];
    inv_aug10[0] = b.inv_augmented_mass.linear;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) inv_aug10[1 + r * 3 + c] = b.inv_augmented_mass.angular.m[r][c];
    acc6[0] = b.acceleration.lin.x; acc6[1] = b.acceleration.lin.y; acc6[2] = b.acceleration.lin.z;
    acc6[3] = b.acceleration.ang.x; acc6[4] = b.acceleration.ang.y; acc6[5] = b.acceleration.ang.z;
    com3[0] = b.com.x; com3[1] = b.com.y; com3[2] = b.com.z;
    return NB2_OK;
}
/* Number of velocity rows per bucket after the last assembly:
 * [joint bilateral, joint bilateral_ground, contact bilateral, contact bilateral_ground,
 *  contact unilateral, contact unilateral_ground]. */
int nbo_debug_row_counts(void* wp, uint32_t* out6) {
    World* w = (World*)wp;
    out6[0] = (uint32_t)w->joint_vel.bilateral.size();
    out6[1] = (uint32_t)w->joint_vel.bilateral_ground.size();
    out6[2] = (uint32_t)w->contact_vel.bilateral.size();
    out6[3] = (uint32_t)w->contact_vel.bilateral_ground.size();
    out6[4] = (uint32_t)w->contact_vel.unilateral.size();
    out6[5] = (uint32_t)w->contact_vel.unilateral_ground.size();
    return NB2_OK;
}
/* mj_lambda_vel of body i after the last velocity solve (zeros for non-dynamic). */
int nbo_debug_mj_lambda(void* wp, uint32_t i, float* out6) {
    World* w = (World*)wp;
    if (i >= w->bodies.size()) return NB2_ERR_BAD_INDEX;
    const Body& b = w->bodies[i];
    for (int k = 0; k < 6; ++k)
        out6[k] = b.status == NB2_BODY_DYNAMIC ? (float)w->mj_lambda_vel[b.companion_id + k] : 0.f;
    return NB2_OK;
}

}  /* extern "C" */
