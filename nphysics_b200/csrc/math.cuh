// Device-side 3D algebra for the solver kernels.
//
// The whole library is compiled with -fmad=false, so every expression below is
// evaluated as written (IEEE mul/add/div/sqrt, round-to-nearest, no FMA
// contraction).  The operation order follows the reference's nalgebra
// expressions (see DESIGN.md "Arithmetic order") so that reference-order mode
// reproduces a sequential f32 evaluation to the last few ulps.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace nb2 {

#define NB2_HD __host__ __device__ __forceinline__
#define NB2_D __device__ __forceinline__

struct Vec3 {
    float x, y, z;
};
NB2_HD Vec3 mk3(float x, float y, float z) {
    Vec3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
NB2_HD Vec3 operator+(Vec3 a, Vec3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
NB2_HD Vec3 operator-(Vec3 a, Vec3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
NB2_HD Vec3 operator-(Vec3 a) { return mk3(-a.x, -a.y, -a.z); }
NB2_HD Vec3 operator*(Vec3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
NB2_HD Vec3 operator/(Vec3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
NB2_HD float dot3(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NB2_HD Vec3 cross3(Vec3 a, Vec3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
NB2_HD float norm_sq3(Vec3 a) { return dot3(a, a); }
NB2_HD float norm3(Vec3 a) { return sqrtf(norm_sq3(a)); }
NB2_HD Vec3 mul3(Vec3 a, Vec3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }

// Row-major 3x3.
struct Mat3 {
    float m[3][3];
};
NB2_HD Vec3 mat_vec(const Mat3& a, Vec3 v) {
    return mk3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
               a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
               a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
NB2_HD Mat3 mat_mul(const Mat3& a, const Mat3& b) {
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
NB2_HD Mat3 mat_transpose(const Mat3& a) {
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
NB2_HD Mat3 cross_matrix(Vec3 v) {
    Mat3 r;
    r.m[0][0] = 0.f;   r.m[0][1] = -v.z;  r.m[0][2] = v.y;
    r.m[1][0] = v.z;   r.m[1][1] = 0.f;   r.m[1][2] = -v.x;
    r.m[2][0] = -v.y;  r.m[2][1] = v.x;   r.m[2][2] = 0.f;
    return r;
}
// Closed-form cofactor inverse; the zero matrix when det == 0 exactly
// (Inertia3::inverse substitutes zero, src/algebra/inertia3.rs:84).
NB2_HD Mat3 mat_inverse_or_zero(const Mat3& a) {
    const float m11 = a.m[0][0], m12 = a.m[0][1], m13 = a.m[0][2];
    const float m21 = a.m[1][0], m22 = a.m[1][1], m23 = a.m[1][2];
    const float m31 = a.m[2][0], m32 = a.m[2][1], m33 = a.m[2][2];
    const float c1 = m22 * m33 - m32 * m23;
    const float c2 = m21 * m33 - m31 * m23;
    const float c3 = m21 * m32 - m31 * m22;
    const float det = m11 * c1 - m12 * c2 + m13 * c3;
    Mat3 r;
    if (det == 0.f) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) r.m[i][j] = 0.f;
        return r;
    }
    r.m[0][0] = c1 / det;
    r.m[0][1] = (m13 * m32 - m33 * m12) / det;
    r.m[0][2] = (m12 * m23 - m22 * m13) / det;
    r.m[1][0] = -c2 / det;
    r.m[1][1] = (m11 * m33 - m31 * m13) / det;
    r.m[1][2] = (m13 * m21 - m23 * m11) / det;
    r.m[2][0] = c3 / det;
    r.m[2][1] = (m12 * m31 - m32 * m11) / det;
    r.m[2][2] = (m11 * m22 - m21 * m12) / det;
    return r;
}

// Unit quaternion (i, j, k, w).
struct Quat {
    float i, j, k, w;
};
NB2_HD Quat mkq(float i, float j, float k, float w) {
    Quat q;
    q.i = i;
    q.j = j;
    q.k = k;
    q.w = w;
    return q;
}
NB2_HD Quat quat_mul(Quat a, Quat b) {
    Quat r;
    r.w = a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k;
    r.i = a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j;
    r.j = a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i;
    r.k = a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w;
    return r;
}
NB2_HD Quat quat_conj(Quat q) { return mkq(-q.i, -q.j, -q.k, q.w); }
NB2_HD Vec3 quat_imag(Quat q) { return mk3(q.i, q.j, q.k); }
NB2_HD Vec3 quat_rotate(Quat q, Vec3 p) {
    Vec3 t = cross3(quat_imag(q), p) * 2.f;
    Vec3 c = cross3(quat_imag(q), t);
    return t * q.w + c + p;
}
NB2_HD Vec3 quat_inv_rotate(Quat q, Vec3 p) { return quat_rotate(quat_conj(q), p); }
NB2_HD Mat3 quat_to_matrix(Quat q) {
    const float i = q.i, j = q.j, k = q.k, w = q.w;
    const float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    const float ij = i * j * 2.f, wk = w * k * 2.f, wj = w * j * 2.f;
    const float ik = i * k * 2.f, jk = j * k * 2.f, wi = w * i * 2.f;
    Mat3 r;
    r.m[0][0] = ww + ii - jj - kk;  r.m[0][1] = ij - wk;            r.m[0][2] = wj + ik;
    r.m[1][0] = wk + ij;            r.m[1][1] = ww - ii + jj - kk;  r.m[1][2] = jk - wi;
    r.m[2][0] = ik - wj;            r.m[2][1] = wi + jk;            r.m[2][2] = ww - ii - jj + kk;
    return r;
}
#define NB2_F32_EPS 1.1920928955078125e-07f
#define NB2_F32_MAX 3.402823466e+38f
#define NB2_PI 3.14159265358979323846f
// exp of the pure quaternion (0, w/2); identity below epsilon.
NB2_D Quat quat_from_scaled_axis(Vec3 axisangle) {
    Vec3 h = axisangle / 2.f;
    float nn = norm_sq3(h);
    if (nn <= NB2_F32_EPS * NB2_F32_EPS) return mkq(0.f, 0.f, 0.f, 1.f);
#ifdef NB2_COLOURED_TU
    // coloured-mode kernels (tolerance-judged): sin(n)/n and cos(n) are even functions of n, so for the small
    // rotations of a position correction (|w| <= max_angular_correction = 0.2, n <= 0.1) their Taylor
    // polynomials in n^2 are exact to f32 rounding (next terms < 1e-11) and need no sqrt, division or sincos
    if (nn <= 0.015625f) {
        const float s_over_n = 1.f + nn * (-1.f / 6.f + nn * (1.f / 120.f + nn * (-1.f / 5040.f)));
        const float c = 1.f + nn * (-0.5f + nn * (1.f / 24.f + nn * (-1.f / 720.f)));
        return mkq(h.x * s_over_n, h.y * s_over_n, h.z * s_over_n, c);
    }
#endif
    float n = sqrtf(nn);
    float s, c;
    sincosf(n, &s, &c);
    Vec3 nv = h * (1.f * s / n);
    return mkq(nv.x, nv.y, nv.z, 1.f * c);
}
NB2_D Quat quat_from_axis_angle(Vec3 axis, float angle) {
    float s, c;
    sincosf(angle / 2.f, &s, &c);
    Vec3 v = axis * s;
    return mkq(v.x, v.y, v.z, c);
}
NB2_D bool unit_try_new_and_get(Vec3 v, float eps, Vec3* dir, float* len) {
    float sq = norm_sq3(v);
    if (sq > eps * eps) {
        float n = sqrtf(sq);
        *dir = v / n;
        *len = n;
        return true;
    }
    return false;
}
NB2_D Vec3 quat_scaled_axis(Quat q) {
    Vec3 v = q.w >= 0.f ? quat_imag(q) : -quat_imag(q);
    float n = norm3(v);
    if (n == 0.f) return mk3(0.f, 0.f, 0.f);
    Vec3 axis = v / n;
    float angle = atan2f(norm3(quat_imag(q)), fabsf(q.w)) * 2.f;
    return axis * angle;
}
NB2_D bool quat_rotation_between_axis(Vec3 na, Vec3 nb, Quat* out) {
    Vec3 c = cross3(na, nb);
    Vec3 axis;
    float len;
    if (unit_try_new_and_get(c, NB2_F32_EPS, &axis, &len)) {
        float cs = dot3(na, nb);
        if (cs <= -1.f) return false;
        if (cs >= 1.f) {
            *out = mkq(0.f, 0.f, 0.f, 1.f);
            return true;
        }
        *out = quat_from_axis_angle(axis, acosf(cs));
        return true;
    } else if (dot3(na, nb) < 0.f) {
        return false;
    }
    *out = mkq(0.f, 0.f, 0.f, 1.f);
    return true;
}
// The two tangents nalgebra's orthonormal_subspace_basis hands to its callback.
NB2_D void tangent_basis(Vec3 n, Vec3* t1, Vec3* t2) {
    Vec3 a;
    if (fabsf(n.x) > fabsf(n.y))
        a = mk3(n.z, 0.f, -n.x);
    else
        a = mk3(0.f, -n.z, n.y);
    a = a / norm3(a);
    *t1 = cross3(a, n);
    *t2 = a;
}

struct Pose {
    Vec3 t;
    Quat r;
};
NB2_HD Pose pose_mul(const Pose& a, const Pose& b) {
    Pose r;
    r.t = a.t + quat_rotate(a.r, b.t);
    r.r = quat_mul(a.r, b.r);
    return r;
}
NB2_HD Vec3 pose_point(const Pose& a, Vec3 p) { return quat_rotate(a.r, p) + a.t; }

}  // namespace nb2
