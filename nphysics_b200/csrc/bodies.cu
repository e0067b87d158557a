// Per-body kernels: state (un)packing, dynamics refresh, integration.
//
// refresh_dynamics replaces RigidBody::update_dynamics + update_acceleration and the
// ext_vels fill of assemble_system (src/object/rigid_body.rs:558-619,
// src/solver/moreau_jean_solver.rs:166-174); integrate_bodies replaces
// update_velocities_and_integrate + RigidBody::integrate / apply_displacement
// (moreau_jean_solver.rs:328-347, rigid_body.rs:371-381,467-505).
#include "solver.cuh"

namespace nb2 {

static const int TPB = 256;
static inline unsigned int nblk(size_t n) { return (unsigned int)((n + TPB - 1) / TPB); }

// raw AoS -> live SoA state (after nb2_upload_bodies)
__global__ void k_unpack_bodies(const nb2_body* __restrict__ raw, unsigned int n, PoseQuads pos_t, PoseQuads pos_q,
                                float4* vel, float4* com_im, int* status, int* true_status) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const nb2_body& b = raw[i];
    Pose p;
    p.t = mk3(b.position[0], b.position[1], b.position[2]);
    p.r = mkq(b.position[3], b.position[4], b.position[5], b.position[6]);
    pos_t[i] = xyz_f4(p.t, 0.f);
    pos_q[i] = quat_f4(p.r);
    vel[2 * i] = make_float4(b.velocity[0], b.velocity[1], b.velocity[2], 0.f);
    vel[2 * i + 1] = make_float4(b.velocity[3], b.velocity[4], b.velocity[5], 0.f);
    Vec3 com = pose_point(p, mk3(b.local_com[0], b.local_com[1], b.local_com[2]));
    com_im[i] = xyz_f4(com, 0.f);
    status[i] = (int)b.status;
    true_status[i] = (int)b.status;
}

__global__ void k_unpack_states(const nb2_body_state* __restrict__ in, const nb2_body* __restrict__ raw,
                                unsigned int first, unsigned int n, PoseQuads pos_t, PoseQuads pos_q, float4* vel,
                                float4* com_im) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    unsigned int i = first + k;
    const nb2_body_state& s = in[k];
    Pose p;
    p.t = mk3(s.position[0], s.position[1], s.position[2]);
    p.r = mkq(s.position[3], s.position[4], s.position[5], s.position[6]);
    pos_t[i] = xyz_f4(p.t, 0.f);
    pos_q[i] = quat_f4(p.r);
    vel[2 * i] = make_float4(s.velocity[0], s.velocity[1], s.velocity[2], 0.f);
    vel[2 * i + 1] = make_float4(s.velocity[3], s.velocity[4], s.velocity[5], 0.f);
    const nb2_body& b = raw[i];
    Vec3 com = pose_point(p, mk3(b.local_com[0], b.local_com[1], b.local_com[2]));
    float im = com_im[i].w;
    com_im[i] = xyz_f4(com, im);
}

__global__ void k_pack_states(nb2_body_state* out, unsigned int first, unsigned int n,
                              ConstPoseQuads pos_t, ConstPoseQuads pos_q,
                              const float4* __restrict__ vel) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    unsigned int i = first + k;
    float4 t = pos_t[i], q = pos_q[i], vl = vel[2 * i], va = vel[2 * i + 1];
    nb2_body_state& s = out[k];
    s.position[0] = t.x; s.position[1] = t.y; s.position[2] = t.z;
    s.position[3] = q.x; s.position[4] = q.y; s.position[5] = q.z; s.position[6] = q.w;
    s.velocity[0] = vl.x; s.velocity[1] = vl.y; s.velocity[2] = vl.z;
    s.velocity[3] = va.x; s.velocity[4] = va.y; s.velocity[5] = va.z;
}

// One thread per body.  World inertia, gyroscopic augmented mass and its inverse,
// acceleration, ext_vels = dt * acceleration; clears mj_lambda.
__global__ void k_refresh_dynamics(const nb2_body* __restrict__ raw, unsigned int n, float dt, Vec3 gravity,
                                   ConstPoseQuads pos_q, const float4* __restrict__ vel,
                                   float4* com_im, float4* inv_i, float4* ext, float4* lam) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const nb2_body& b = raw[i];
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    lam[2 * i] = z;
    lam[2 * i + 1] = z;
    float4 c = com_im[i];
    if (b.status != NB2_BODY_DYNAMIC) {
        com_im[i] = make_float4(c.x, c.y, c.z, 0.f);
        inv_i[3 * i] = z;
        inv_i[3 * i + 1] = z;
        inv_i[3 * i + 2] = z;
        ext[2 * i] = z;
        ext[2 * i + 1] = z;
        return;
    }
    // update_dynamics (rigid_body.rs:558-588)
    Mat3 il;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) il.m[r][cc] = b.local_inertia[r * 3 + cc];
    Mat3 rot = quat_to_matrix(f4_quat(pos_q[i]));
    Mat3 iw_ = mat_mul(mat_mul(rot, il), mat_transpose(rot));  // Inertia3::transformed (inertia3.rs:68-71)
    Vec3 w = f4_xyz(vel[2 * i + 1]);
    Vec3 iw = mat_vec(iw_, w);
    Mat3 w_dt_cross = cross_matrix(w * dt);
    Mat3 iw_dt_cross = cross_matrix(iw * dt);
    Mat3 wi = mat_mul(w_dt_cross, iw_);
    Mat3 aug;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) aug.m[r][cc] = iw_.m[r][cc] + (wi.m[r][cc] - iw_dt_cross.m[r][cc]);
    Mat3 inv = mat_inverse_or_zero(aug);
    float inv_mass = b.mass == 0.f ? 0.f : 1.f / b.mass;
    // update_acceleration (rigid_body.rs:590-619)
    Vec3 gyroscopic = -cross3(w, iw);
    Vec3 acc_ang = mat_vec(inv, gyroscopic);
    Vec3 acc_lin = mk3(0.f, 0.f, 0.f);
    if (inv_mass != 0.f && (b.flags & NB2_BODY_FLAG_GRAVITY)) acc_lin = gravity;
    Vec3 f_lin = mk3(b.external_forces[0], b.external_forces[1], b.external_forces[2]);
    Vec3 f_ang = mk3(b.external_forces[3], b.external_forces[4], b.external_forces[5]);
    acc_lin = acc_lin + f_lin * inv_mass;
    acc_ang = acc_ang + mat_vec(inv, f_ang);
    acc_lin = mul3(acc_lin, mk3(b.jacobian_mask[0], b.jacobian_mask[1], b.jacobian_mask[2]));
    acc_ang = mul3(acc_ang, mk3(b.jacobian_mask[3], b.jacobian_mask[4], b.jacobian_mask[5]));
    com_im[i] = make_float4(c.x, c.y, c.z, inv_mass);
    lam[2 * i] = make_float4(0.f, 0.f, 0.f, inv_mass);  // spare lane: the solve kernels rebuild WJ.lin = J.lin * inv_mass
    inv_i[3 * i] = make_float4(inv.m[0][0], inv.m[0][1], inv.m[0][2], 0.f);
    inv_i[3 * i + 1] = make_float4(inv.m[1][0], inv.m[1][1], inv.m[1][2], 0.f);
    inv_i[3 * i + 2] = make_float4(inv.m[2][0], inv.m[2][1], inv.m[2][2], 0.f);
    ext[2 * i] = make_float4(dt * acc_lin.x, dt * acc_lin.y, dt * acc_lin.z, 0.f);
    ext[2 * i + 1] = make_float4(dt * acc_ang.x, dt * acc_ang.y, dt * acc_ang.z, 0.f);
}

// v += ext_vels + mj_lambda; damping; velocity caps; displacement about the com.
__global__ void k_integrate(const nb2_body* __restrict__ raw, unsigned int n, float dt, int kinematic_only,
                            PoseQuads pos_t, PoseQuads pos_q, float4* vel, float4* com_im,
                            const float4* __restrict__ ext, const float4* __restrict__ lam) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const nb2_body& b = raw[i];
    if (kinematic_only ? (b.status != NB2_BODY_KINEMATIC) : (b.status != NB2_BODY_DYNAMIC)) return;
    Vec3 vl = f4_xyz(vel[2 * i]), va = f4_xyz(vel[2 * i + 1]);
    if (!kinematic_only) {
        vl = (vl + f4_xyz(ext[2 * i])) + f4_xyz(lam[2 * i]);
        va = (va + f4_xyz(ext[2 * i + 1])) + f4_xyz(lam[2 * i + 1]);
    }
    vl = vl * (1.f / (1.f + dt * b.linear_damping));
    va = va * (1.f / (1.f + dt * b.angular_damping));
    float ln = norm3(vl);
    if (ln > b.max_linear_velocity) {
        if (b.max_linear_velocity == 0.f)
            vl = mk3(0.f, 0.f, 0.f);
        else
            vl = vl * (b.max_linear_velocity / ln);
    }
    float an = norm3(va);
    if (an > b.max_angular_velocity) {
        if (b.max_angular_velocity == 0.f)
            va = mk3(0.f, 0.f, 0.f);
        else
            va = va * (b.max_angular_velocity / an);
    }
    BodyPose bp;
    bp.pose.t = f4_xyz(pos_t[i]);
    bp.pose.r = f4_quat(pos_q[i]);
    float4 c = com_im[i];
    bp.com = f4_xyz(c);
    apply_displacement(&bp, mk3(b.local_com[0], b.local_com[1], b.local_com[2]), vl * dt, va * dt);
    pos_t[i] = xyz_f4(bp.pose.t, 0.f);
    pos_q[i] = quat_f4(bp.pose.r);
    com_im[i] = xyz_f4(bp.com, c.w);
    vel[2 * i] = xyz_f4(vl, 0.f);
    vel[2 * i + 1] = xyz_f4(va, 0.f);
}

// kinetic energy + non-finite guard
__global__ void k_body_stats(const nb2_body* __restrict__ raw, unsigned int n, ConstPoseQuads pos_t,
                             ConstPoseQuads pos_q, const float4* __restrict__ vel, double* energy,
                             unsigned int* non_finite) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    double ke = 0.0;
    unsigned int bad = 0;
    if (i < n && raw[i].status == NB2_BODY_DYNAMIC) {
        const nb2_body& b = raw[i];
        Mat3 il;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) il.m[r][cc] = b.local_inertia[r * 3 + cc];
        float4 q = pos_q[i], t = pos_t[i];
        Mat3 rot = quat_to_matrix(f4_quat(q));
        Mat3 iw_ = mat_mul(mat_mul(rot, il), mat_transpose(rot));
        Vec3 vl = f4_xyz(vel[2 * i]), va = f4_xyz(vel[2 * i + 1]);
        Vec3 l = mat_vec(iw_, va);
        ke = 0.5 * (double)b.mass * (double)norm_sq3(vl) + 0.5 * (double)dot3(va, l);
        float chk = t.x + t.y + t.z + q.x + q.y + q.z + q.w + vl.x + vl.y + vl.z + va.x + va.y + va.z;
        bad = isfinite(chk) ? 0u : 1u;
    }
    // warp reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) {
        ke += __shfl_down_sync(0xffffffffu, ke, o);
        bad += __shfl_down_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (ke != 0.0) atomicAdd(energy, ke);
        if (bad) atomicAdd(non_finite, bad);
    }
}

int launch_unpack_bodies(Context* ctx) {
    if (!ctx->n_bodies) return NB2_OK;
    k_unpack_bodies<<<nblk(ctx->n_bodies), TPB, 0, ctx->stream>>>(ctx->raw.p, ctx->n_bodies, ctx->pos_t.p,
                                                                  ctx->pos_q.p, ctx->vel.p, ctx->com_im.p,
                                                                  ctx->b_status.p, ctx->true_status.p);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int launch_unpack_states(Context* ctx, const nb2_body_state* d_in, uint32_t first, uint32_t n) {
    if (!n) return NB2_OK;
    k_unpack_states<<<nblk(n), TPB, 0, ctx->stream>>>(d_in, ctx->raw.p, first, n, ctx->pos_t.p, ctx->pos_q.p,
                                                      ctx->vel.p, ctx->com_im.p);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int launch_pack_states(Context* ctx, nb2_body_state* d_out, uint32_t first, uint32_t n) {
    if (!n) return NB2_OK;
    k_pack_states<<<nblk(n), TPB, 0, ctx->stream>>>(d_out, first, n, ctx->pos_t.p, ctx->pos_q.p, ctx->vel.p);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int launch_refresh_dynamics(Context* ctx) {
    Vec3 g = mk3(ctx->params.gravity[0], ctx->params.gravity[1], ctx->params.gravity[2]);
    k_refresh_dynamics<<<nblk(ctx->n_bodies), TPB, 0, ctx->stream>>>(ctx->raw.p, ctx->n_bodies, ctx->params.dt, g,
                                                                     ctx->pos_q.p, ctx->vel.p, ctx->com_im.p,
                                                                     ctx->inv_i.p, ctx->ext.p, ctx->lam.p);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int launch_integrate(Context* ctx, bool kinematic_only) {
    k_integrate<<<nblk(ctx->n_bodies), TPB, 0, ctx->stream>>>(ctx->raw.p, ctx->n_bodies, ctx->params.dt,
                                                              kinematic_only ? 1 : 0, ctx->pos_t.p, ctx->pos_q.p,
                                                              ctx->vel.p, ctx->com_im.p, ctx->ext.p, ctx->lam.p);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int launch_body_stats(Context* ctx, double* d_energy, unsigned int* d_non_finite) {
    k_body_stats<<<nblk(ctx->n_bodies), TPB, 0, ctx->stream>>>(ctx->raw.p, ctx->n_bodies, ctx->pos_t.p, ctx->pos_q.p,
                                                               ctx->vel.p, d_energy, d_non_finite);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

}  // namespace nb2
