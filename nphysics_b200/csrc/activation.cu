// Sleeping: ActivationManager::update (src/detection/activation_manager.rs:60-201) on device.
//
// The reference walks the bodies, a union-find over the contact pairs and joints, and two more passes
// over the bodies, all sequentially.  Here: one kernel for the energy low-pass, a cooperative kernel
// that labels the islands by min-label hooking + pointer jumping (the partition is the one the
// reference's union-find produces; only the representative differs, and nothing depends on it), one
// kernel for the per-island verdict and one that puts bodies to sleep / wakes them.
//
// A sleeping dynamic body is then made invisible to the step the same way the reference does it
// (it is not in active_bodies, and pairs without an awake dynamic body are not in the manifold list,
// mechanical_world.rs:287-300): its EFFECTIVE status -- raw[].status and b_status[], the only status
// the step kernels read -- becomes static, while true_status[] keeps what the caller uploaded.
#include "solver.cuh"

namespace nb2 {

static const int TPB = 256;
static inline unsigned int nblk(size_t n) { return (unsigned int)((n + TPB - 1) / TPB); }

__device__ __forceinline__ bool act_eligible(int st) {  // status_dependent_ndofs() != 0 || is_kinematic()
    return st == NB2_BODY_DYNAMIC || st == NB2_BODY_KINEMATIC;
}

// update_energy (activation_manager.rs:47-58) for the awake dynamic bodies; also resets the labels
__global__ void k_act_energy(unsigned int n, const int* __restrict__ true_status, const float4* __restrict__ vel,
                             float2* act, unsigned int* parent, unsigned int* can, float mix) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    parent[i] = i;
    can[i] = 1u;
    if (true_status[i] != NB2_BODY_DYNAMIC) return;
    float2 a = act[i];
    if (a.y == 0.f || a.x < 0.f) return;
    const float4 vl = vel[2 * i], va = vel[2 * i + 1];
    const float v[6] = {vl.x, vl.y, vl.z, va.x, va.y, va.z};
    float nsq = 0.f;  // nalgebra norm_squared of a 6-slice: sequential
#pragma unroll
    for (int k = 0; k < 6; ++k) nsq += v[k] * v[k];
    const float e = (1.f - mix) * a.y + mix * nsq;
    act[i] = make_float2(a.x, fminf(e, a.x * 4.f));
}

// deferred_activate handles (activation_manager.rs:100-108; Body::activate, body.rs:338-342)
__global__ void k_act_wake(unsigned int n_list, const int* __restrict__ list, unsigned int n, float2* act) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_list) return;
    const int i = list[k];
    if (i < 0 || (unsigned int)i >= n) return;
    const float thr = act[i].x;
    if (thr >= 0.f) act[i].y = thr * 2.f;  // duplicates write the same value
}

// Islands: connected components over the eligible bodies, edges = manifolds with contacts and unbroken
// joints (make_union, activation_manager.rs:138-166).  parent[] converges to the smallest body index
// of the component.
__global__ void __launch_bounds__(TPB) k_act_islands(unsigned int nb, unsigned int nm, unsigned int nj,
                                                     const nb2_manifold* __restrict__ manifolds,
                                                     const nb2_joint* __restrict__ joints,
                                                     const int* __restrict__ true_status, unsigned int* parent,
                                                     unsigned int* flags /*3*/, unsigned int* barrier, int dynamic_only) {
    GridBarrier gb;
    gb.init(barrier);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (unsigned int round = 1;; ++round) {
        unsigned int changed = 0;
        for (size_t e = tid; e < (size_t)nm + nj; e += stride) {
            int b1, b2;
            if (e < nm) {
                const nb2_manifold& m = manifolds[e];
                if (m.num_contacts == 0 && !dynamic_only) continue;  // sharding keeps potential pairs together
                b1 = m.body1;
                b2 = m.body2;
            } else {
                const nb2_joint& j = joints[e - nm];
                if (j.broken) continue;
                b1 = j.body1;
                b2 = j.body2;
            }
            if ((unsigned int)b1 >= nb || (unsigned int)b2 >= nb) continue;  // bad records are reported by the step
            if (dynamic_only ? (true_status[b1] != NB2_BODY_DYNAMIC || true_status[b2] != NB2_BODY_DYNAMIC)
                             : (!act_eligible(true_status[b1]) || !act_eligible(true_status[b2])))
                continue;
            // hook the larger root under the smaller
            unsigned int r1 = __ldcg(&parent[b1]), r2 = __ldcg(&parent[b2]);
            while (r1 != r2) {
                const unsigned int hi = max(r1, r2), lo = min(r1, r2);
                const unsigned int old = atomicMin(&parent[hi], lo);
                changed = 1;
                if (old == hi || old == lo) break;
                r1 = old;  // someone else re-parented `hi`: continue from there
                r2 = lo;
            }
        }
        if (__syncthreads_or(changed) && threadIdx.x == 0) atomicOr(&flags[round % 3], 1u);
        gb.sync();
        // pointer jumping to the root
        for (size_t i = tid; i < nb; i += stride) {
            unsigned int p = __ldcg(&parent[i]);
            unsigned int g = __ldcg(&parent[p]);
            while (p != g) {
                p = g;
                g = __ldcg(&parent[p]);
            }
            __stcg(&parent[i], p);
        }
        const unsigned int any = *((volatile unsigned int*)&flags[round % 3]);
        if (tid == 0) flags[(round + 2) % 3] = 0;
        gb.sync();
        if (!any) break;
    }
}

// can_deactivate[root] (activation_manager.rs:170-182)
__global__ void k_act_verdict(unsigned int n, const int* __restrict__ true_status, const float2* __restrict__ act,
                              const unsigned int* __restrict__ parent, unsigned int* can) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !act_eligible(true_status[i])) return;
    const float2 a = act[i];
    if (!(a.x >= 0.f && a.y < a.x)) can[parent[i]] = 0u;
}

// put to sleep / wake up (activation_manager.rs:185-206), then refresh the effective status
__global__ void k_act_apply(unsigned int n, const int* __restrict__ true_status, float2* act,
                            const unsigned int* __restrict__ parent, const unsigned int* __restrict__ can, float4* vel,
                            nb2_body* raw, int* b_status, int run_manager) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int st = true_status[i];
    float2 a = act[i];
    if (run_manager && act_eligible(st)) {
        if (can[parent[i]]) {
            if (a.y != 0.f) {  // RigidBody::deactivate (rigid_body.rs:396-400)
                a.y = 0.f;
                vel[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
                vel[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else if (st != NB2_BODY_KINEMATIC) {
            if (a.y == 0.f && a.x >= 0.f) a.y = a.x * 2.f;
        }
        act[i] = a;
    }
    const int eff = (st == NB2_BODY_DYNAMIC && a.y == 0.f) ? NB2_BODY_STATIC : st;
    raw[i].status = (uint32_t)eff;
    b_status[i] = eff;
}

int launch_apply_effective_status(Context* ctx) {
    if (!ctx->sleeping || !ctx->n_bodies) return NB2_OK;
    k_act_apply<<<nblk(ctx->n_bodies), TPB, 0, ctx->stream>>>(ctx->n_bodies, ctx->true_status.p, ctx->act.p, nullptr, nullptr,
                                                              ctx->vel.p, ctx->raw.p, ctx->b_status.p, 0);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

static int launch_islands(Context* ctx, int dynamic_only);

// nb2_label_islands: the sharding view of the same labelling -- connected components over the
// DYNAMIC bodies only (static and kinematic bodies are replicated into every shard, so they must not glue
// islands together: activation_manager.rs:141-145), plus the velocity rows each body owns (a group's rows
// are booked on its first dynamic body), from which the host bin-packs islands onto ranks.
__global__ void k_island_init(unsigned int n, unsigned int* parent, unsigned int* rows) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    parent[i] = i;
    rows[i] = 0u;
}
__global__ void k_island_rows(unsigned int nb, unsigned int nm, unsigned int nj, const nb2_manifold* __restrict__ manifolds,
                              const nb2_joint* __restrict__ joints, const int* __restrict__ true_status, unsigned int* rows) {
    unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nm + nj) return;
    int b1, b2;
    unsigned int r;
    if (e < nm) {
        b1 = manifolds[e].body1;
        b2 = manifolds[e].body2;
        r = 3u * manifolds[e].num_contacts;
    } else {
        const nb2_joint& j = joints[e - nm];
        if (j.broken) return;
        const unsigned int per_type[NB2_JOINT_TYPE_COUNT] = {3, 5, 5, 4, 3, 4, 4, 4, 6, 3};
        b1 = j.body1;
        b2 = j.body2;
        r = j.type < NB2_JOINT_TYPE_COUNT ? per_type[j.type] : 0u;
    }
    if ((unsigned int)b1 >= nb || (unsigned int)b2 >= nb || r == 0u) return;
    const int owner = true_status[b1] == NB2_BODY_DYNAMIC ? b1 : (true_status[b2] == NB2_BODY_DYNAMIC ? b2 : -1);
    if (owner >= 0) atomicAdd(&rows[owner], r);
}
__global__ void k_island_labels(unsigned int n, const int* __restrict__ true_status, const unsigned int* __restrict__ parent,
                                int* labels) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    labels[i] = true_status[i] == NB2_BODY_DYNAMIC ? (int)parent[i] : -1;
}

int launch_label_islands(Context* ctx, int* d_labels, unsigned int* d_rows) {
    const unsigned int nb = ctx->n_bodies;
    NB2_TRY(ctx->cc_parent.reserve(ctx, (size_t)nb + 1));
    k_island_init<<<nblk(nb), TPB, 0, ctx->stream>>>(nb, ctx->cc_parent.p, d_rows);
    ctx->launches++;
    const unsigned int ne = ctx->n_manifolds + ctx->n_joints;
    if (ne) {
        NB2_TRY(launch_islands(ctx, 1));
        k_island_rows<<<nblk(ne), TPB, 0, ctx->stream>>>(nb, ctx->n_manifolds, ctx->n_joints, ctx->manifolds.p, ctx->joints.p,
                                                        ctx->true_status.p, d_rows);
        ctx->launches++;
    }
    k_island_labels<<<nblk(nb), TPB, 0, ctx->stream>>>(nb, ctx->true_status.p, ctx->cc_parent.p, d_labels);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int launch_update_activation(Context* ctx, float mix, const int32_t* to_activate, uint32_t n_list) {
    const unsigned int nb = ctx->n_bodies;
    NB2_TRY(ctx->cc_parent.reserve(ctx, (size_t)nb + 1));
    NB2_TRY(ctx->cc_can.reserve(ctx, (size_t)nb + 1));
    k_act_energy<<<nblk(nb), TPB, 0, ctx->stream>>>(nb, ctx->true_status.p, ctx->vel.p, ctx->act.p, ctx->cc_parent.p,
                                                   ctx->cc_can.p, mix);
    ctx->launches++;
    if (n_list) {
        NB2_TRY(ctx->wake_list.reserve(ctx, n_list));
        NB2_CUDA(ctx, cudaMemcpyAsync(ctx->wake_list.p, to_activate, (size_t)n_list * sizeof(int32_t), cudaMemcpyHostToDevice,
                                      ctx->stream));
        NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller may free the list on return
        k_act_wake<<<nblk(n_list), TPB, 0, ctx->stream>>>(n_list, ctx->wake_list.p, nb, ctx->act.p);
        ctx->launches++;
    }
    if (ctx->n_manifolds + ctx->n_joints) NB2_TRY(launch_islands(ctx, 0));
    k_act_verdict<<<nblk(nb), TPB, 0, ctx->stream>>>(nb, ctx->true_status.p, ctx->act.p, ctx->cc_parent.p, ctx->cc_can.p);
    k_act_apply<<<nblk(nb), TPB, 0, ctx->stream>>>(nb, ctx->true_status.p, ctx->act.p, ctx->cc_parent.p, ctx->cc_can.p,
                                                   ctx->vel.p, ctx->raw.p, ctx->b_status.p, 1);
    ctx->launches += 2;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}


// the cooperative labelling kernel over the uploaded / produced manifolds and the joints
static int launch_islands(Context* ctx, int dynamic_only) {
    const unsigned int nb = ctx->n_bodies;
    {
        NB2_TRY(ctx->barrier.reserve(ctx, NB2_BARRIER_WORDS));
        NB2_CUDA(ctx, cudaMemsetAsync(ctx->barrier.p, 0, 8 * sizeof(unsigned int), ctx->stream));
        int& blocks_cc = ctx->coop_blocks_islands;
        if (blocks_cc <= 0) {
            int per_sm = 0;
            NB2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_act_islands, TPB, 0));
            if (per_sm < 1) return set_error(ctx, NB2_ERR_CUDA, "cooperative kernel does not fit on an SM");
            blocks_cc = (per_sm > 4 ? 4 : per_sm) * ctx->sm_count;
        }
        const size_t work = (size_t)ctx->n_manifolds + ctx->n_joints > nb ? (size_t)ctx->n_manifolds + ctx->n_joints : nb;
        int blocks = (int)((work + TPB - 1) / TPB);
        if (blocks > blocks_cc) blocks = blocks_cc;
        if (blocks < 1) blocks = 1;
        unsigned int nb_ = nb, nm = ctx->n_manifolds, nj = ctx->n_joints;
        const nb2_manifold* mf = ctx->manifolds.p;
        const nb2_joint* jt = ctx->joints.p;
        const int* ts = ctx->true_status.p;
        unsigned int* parent = ctx->cc_parent.p;
        unsigned int* flags = ctx->barrier.p + 4;
        unsigned int* bar = ctx->barrier.p;
        void* args[] = {&nb_, &nm, &nj, &mf, &jt, &ts, &parent, &flags, &bar, &dynamic_only};
        NB2_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_act_islands, dim3(blocks), dim3(TPB), args, 0, ctx->stream));
        ctx->launches++;
    }
    return NB2_OK;
}

}  // namespace nb2
