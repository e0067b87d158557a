// Internal definitions shared by the .cu translation units of libnphysics_b200.so.
//
// Layout of the step (DESIGN.md has the full picture):
//   refresh_dynamics -> schedule (levels | colours) -> assemble rows -> PGS velocity
//   -> cache impulses -> integrate -> PGS position -> kinematic integrate
//
// All hot data are SoA float4 streams indexed by "slot".  Rows of one group
// (the rows of <= 4 contacts of one manifold, or the rows of one joint) are
// strided by the number of groups in their phase (ELL layout) so that thread g
// of a phase reads row r of its group at  rbase + r*count + g  -- consecutive
// threads touch consecutive 16-byte words.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nphysics_b200.h"
#include "math.cuh"

namespace nb2 {

// the kernels read parts of the body record as whole quads (assemble.cu load_side, solve_position.cuh)
static_assert(sizeof(nb2_body) == 176 && offsetof(nb2_body, local_com) == 52 && offsetof(nb2_body, jacobian_mask) == 144 &&
                  offsetof(nb2_body, status) == 168,
              "nb2_body layout changed: update the quad loads");
static_assert(sizeof(nb2_contact) == 112, "nb2_contact layout changed: update the quad loads in assemble.cu");

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
struct Context;
int set_error(Context* ctx, int code, const char* fmt, ...);

#define NB2_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return set_error((ctx), e_ == cudaErrorMemoryAllocation ? NB2_ERR_OUT_OF_MEMORY      \
                                                                    : NB2_ERR_CUDA,              \
                             "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,   \
                             __LINE__);                                                          \
    } while (0)

#define NB2_TRY(expr)             \
    do {                          \
        int rc_ = (expr);         \
        if (rc_ != NB2_OK) return rc_; \
    } while (0)

// Growable device buffer (context-owned).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    int reserve(Context* ctx, size_t n) {
        if (n <= cap) return NB2_OK;
        size_t ncap = n + n / 4 + 64;
        T* np_ = nullptr;
        cudaError_t e = cudaMalloc((void**)&np_, ncap * sizeof(T));
        if (e != cudaSuccess)
            return set_error(ctx, NB2_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu bytes) failed: %s", ncap * sizeof(T),
                             cudaGetErrorString(e));
        if (p) cudaFree(p);  // contents are never carried over
        p = np_;
        cap = ncap;
        return NB2_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Strided view of the interleaved pose buffer: element i is quad 2 * i (+ 1 for the rotation).
struct ConstPoseQuads {
    const float4* p;
    __host__ __device__ __forceinline__ const float4& operator[](size_t i) const { return p[2 * i]; }
};
struct PoseQuads {
    float4* p;
    __host__ __device__ __forceinline__ float4& operator[](size_t i) const { return p[2 * i]; }
    __host__ __device__ __forceinline__ operator ConstPoseQuads() const { return ConstPoseQuads{p}; }
};

// ---------------------------------------------------------------------------
// scheduling
// ---------------------------------------------------------------------------
#define NB2_CHUNK 4            // contacts per contact group
#define NB2_MAX_JOINT_ROWS 7   // prismatic reserves 7 (prismatic_constraint.rs:124-126)
#define NB2_MASK_WORDS 4       // 4 x 64 colours
#define NB2_MAX_COLOURS (64 * NB2_MASK_WORDS)
#define NB2_ROW_PLANES 5       // jacobian quads per velocity row (solve_common.cuh)
#define NB2_STAGED_NOT_APPLICABLE 1  // launch_position_solve_staged: not an error (those are negative), use the plain kernel
#ifndef NB2_BALANCE_ROUNDS
#define NB2_BALANCE_ROUNDS 12
#endif
#ifndef NB2_IG_PASSES
#define NB2_IG_PASSES 1        // iterated-greedy passes right after a fresh colouring ...
#endif
#define NB2_IG_REFINE 12       // ... and one more per following step with an unchanged conflict graph, up to this many

// velocity-row kinds
#define NB2_ROW_NONE 0
#define NB2_ROW_UNILATERAL 1
#define NB2_ROW_BILATERAL 2   // Independent{lo, hi}
#define NB2_ROW_DEPENDENT 3   // Dependent{dependency, coeff}

// g_info.z low byte: generic row count (row layout, low nibble != 0 for contact groups: 3,6,9,12) or
// (contacts << 4) for compact contact groups
#define NB2_Z_IS_COMPACT(z) ((((z) >> 8) == NB2_ITEM_CONTACTS) && (((z) & 0xF) == 0))

// item types
#define NB2_ITEM_JOINT 0
#define NB2_ITEM_CONTACTS 1     // coloured: friction + normal rows of one chunk
#define NB2_ITEM_FRICTION 2     // reference-order: the friction rows of one chunk
#define NB2_ITEM_NORMAL 3       // reference-order: the normal rows of one chunk
#define NB2_ITEM_INVALID -1

// Device-resident header of a schedule (read by kernels; never read back by the
// host on the step path).
struct SchedHeader {
    unsigned int n_phases;
    unsigned int n_groups;   // total scheduled items
    unsigned int n_slots;    // total row slots
    unsigned int overflow;   // != 0: colouring ran out of colours / phases
    unsigned int work;       // velocity rows actually scheduled (n_slots counts the ELL padding too)
    unsigned int refine_left;  // iterated-greedy passes still to run on steps whose conflict graph is unchanged
    unsigned int pad[2];       // [0] incremental recolourings since the last colouring from scratch, [1] groups of the largest phase
};

struct Sched {
    size_t n_items = 0;      // capacity in items for this step
    size_t max_phases = 0;
    DevBuf<int> it_a, it_b;          // dynamic body of each side or -1
    DevBuf<int> it_nrows;            // velocity rows (position schedule: position constraints)
    DevBuf<int> it_type;             // NB2_ITEM_*
    DevBuf<int> it_src;              // joint index / chunk index
    DevBuf<unsigned long long> it_key;  // sequential position (reference order)
    DevBuf<int> it_phase, it_slot;   // outputs: phase, index inside the phase
    DevBuf<int> it_phase_raw;        // coloured: the colouring before balancing (what the refinement passes work on)
    DevBuf<unsigned int> ph_count, ph_R, ph_gbase, ph_rbase;
    DevBuf<unsigned int> ph_bcnt;    // [phase][row-count bucket]: counts, then slot cursors
    DevBuf<int4> g_info;             // per group slot: (a, b, nrows | type << 8, item)
    DevBuf<SchedHeader> hdr;         // 1 element
    // schedule cache: last step's groups, compared on device (coloured mode)
    DevBuf<int> prev_a, prev_b, prev_nt, prev_b1, prev_b2;
    DevBuf<int> it_b1, it_b2;        // body indices of both sides, whatever their status
    bool cache_valid = false;
    size_t cache_n = 0;
    unsigned int cache_bodies = 0;
    void release() {
        it_a.release(); it_b.release(); it_nrows.release(); it_type.release(); it_src.release();
        it_key.release(); it_phase.release(); it_slot.release(); ph_count.release(); ph_R.release();
        ph_gbase.release(); ph_rbase.release(); g_info.release(); hdr.release();
        prev_a.release(); prev_b.release(); prev_nt.release(); prev_b1.release(); prev_b2.release();
        it_b1.release(); it_b2.release(); it_phase_raw.release(); ph_bcnt.release();
    }
};

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct StageEvents {
    // 0 step start, 1 assembly done, 2 velocity done, 3 update done, 4 position done, 5 end,
    // 6/7 around the velocity kernel, 8/9 around the position kernel, 10/11 around scheduling
    cudaEvent_t e[12];
    bool created = false;
};

struct Context {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    char last_error[512];
    uint64_t launches = 0;

    nb2_params params;
    float inv_dt = 0.f;
    bool have_params = false;
    bool timers = false;
    bool schedule_cache = true;
    bool incremental_colouring = true;  // NB2_INCREMENTAL_COLOURING=0: any change of the conflict graph colours from scratch
    bool manifolds_from_producer = false;  // the contact set of the next step was written by nb2_generate_manifolds
    // coloured mode, contact groups: 0 = 132-byte row stream (default), 1 = 80-byte compact records
    int contact_model = 0;  // nb2_contact_model
    int contact_layout = 0;
    int step_layout = 0;  // layout the last assembly used
    StageEvents ev;
    bool ev_valid = false;

    // ---- bodies
    uint32_t n_bodies = 0;
    uint32_t n_dynamic = 0;
    unsigned int n_kinematic = 0;  // KINEMATIC bodies of the uploaded set (none: the end-of-step pass is not launched)
    DevBuf<nb2_body> raw;         // static properties (pose/velocity fields are upload-time values)
    // live pose: (t.xyz, 0) and the rotation quaternion of a body are neighbours in ONE buffer, so a pose is one
    // 32-byte sector (the position solve gathers two poses per group visit and is bound by the sectors it pulls
    // through the SM's L2 port, profiles/r02_notes.md); pos_t.p / pos_q.p are strided views of it
    DevBuf<float4> pos;
    struct PoseView { PoseQuads p; } pos_t, pos_q;
    DevBuf<float4> vel;           // [2n] linear, angular
    DevBuf<float4> com_im;        // com.xyz, inverse mass
    DevBuf<float4> inv_i;         // [3n] rows of the inverse augmented angular inertia
    DevBuf<float4> ext;           // [2n] ext_vels = dt * acceleration
    DevBuf<float4> lam;           // [2n] mj_lambda_vel
    DevBuf<int> b_status;            // EFFECTIVE status (a sleeping dynamic body reads as static)
    // ---- sleeping (activation.cu); off until nb2_upload_activation
    bool sleeping = false;
    DevBuf<int> true_status;         // status as uploaded
    DevBuf<float2> act;              // threshold (< 0: None), energy (0: asleep)
    DevBuf<unsigned int> cc_parent, cc_can;
    DevBuf<int> wake_list;
    DevBuf<int> isl_labels;          // nb2_label_islands

    // ---- joints
    uint32_t n_joints = 0;
    DevBuf<nb2_joint> joints;

    // ---- contacts
    uint32_t n_manifolds = 0, n_contacts = 0;
    DevBuf<nb2_manifold> manifolds;
    DevBuf<nb2_contact> contacts;
    DevBuf<unsigned int> c_manifold;   // contact -> manifold
    DevBuf<unsigned int> chunk_base;   // [nM+1] exclusive scan of ceil(nc/4)
    DevBuf<unsigned int> chunk_manifold;  // chunk -> manifold
    size_t max_chunks = 0;

    // impulse cache: double-buffered (previous step -> this step)
    DevBuf<float4> imp[2];
    DevBuf<unsigned long long> ckey[2];  // key of every contact of that step (the cache's fast path: same key at the same index)
    DevBuf<unsigned long long> ht_keys[2];
    DevBuf<float4> ht_imps[2];       // cached impulses (normal, tangent 1, tangent 2) at the key's index
    size_t ht_cap[2] = {0, 0};  // power of two (0 = empty cache)
    uint32_t imp_n[2] = {0, 0};
    int cur = 0;                // buffer written by the current step

    // ---- device manifold producer (narrowphase.cu)
    uint32_t n_colliders = 0, n_pairs = 0;
    bool pairs_valid = false;
    float np_prediction = 0.001f;
    DevBuf<nb2_collider> colliders;
    DevBuf<float4> coll_world;            // collider centre in world space, w = 1 for a dynamic collider
    DevBuf<unsigned int> np_is_big, np_big_off, np_par, pair_cnt, pair_off, grid_count, grid_off, grid_cursor;
    DevBuf<int> np_big_list, grid_entries, pair_feat;
    DevBuf<float4> pair_q;                // [3][cap] persistent feature pairs
    DevBuf<nb2_contact_update> contact_updates;

    // ---- schedules
    Sched vs, ps;          // velocity / position (coloured mode uses vs for both)
    // reference-order scratch
    DevBuf<unsigned int> deg, adj_off, cursor;
    DevBuf<int> adj, pred_a, pred_b, level;
    DevBuf<unsigned int> adj_off_p;  // adjacency of the position schedule (kept apart: the velocity
    DevBuf<int> adj_p;               // adjacency is read again by the reference-order warm start)
    // colouring scratch
    DevBuf<unsigned long long> cmask, best;
    // scan scratch
    DevBuf<unsigned int> scan_tmp;
    // grid barrier counter
    DevBuf<unsigned int> barrier;

    // ---- rows
    size_t n_slots_max = 0, n_pslots_max = 0;
    DevBuf<float4> r_jac;   // [NB2_ROW_PLANES][n_slots_max] (layout: solve_common.cuh)
    DevBuf<float4> r_hdr;   // rhs, r, lo|mu, hi
    DevBuf<float> r_imp;
    DevBuf<float4> p_row;   // [5][n_pslots_max] position rows
    // coloured mode: compact velocity data of a contact, same slot index as p_row:
    // (p1, rhs_n) (p2, rhs_t1) (n, rhs_t2) (r_n, r_t1, r_t2, mu) (imp_n, imp_t1, imp_t2, valid)
    DevBuf<float4> c_geo;   // [5][n_pslots_max]
    bool any_mask = false;  // some body has a jacobian_mask entry != 1

    // ---- stats
    DevBuf<float> stat_f;           // reductions
    DevBuf<unsigned int> stat_u;
    DevBuf<unsigned int> flags;     // [0] input validation bits
    DevBuf<nb2_body_state> stage_states;
    nb2_stats last_stats;
    int last_mode = -1;
    bool stepped = false;

    // co-resident block limits of the cooperative kernels on this context's device (queried once)
    int coop_blocks_vel = 0, coop_blocks_pos = 0, coop_blocks_col = 0, coop_blocks_level = 0, coop_blocks_colour = 0,
        coop_blocks_islands = 0;
    // coloured solve kernels: 0 = phase barrier + register pipelining (the reference-order kernels), 2 = staged
    // (rows streamed through a shared-memory ring with cp.async, prefetched across the phase barrier; the default:
    // free-running rings for a uniform schedule, warp lockstep for a ragged one), 3 = ring filled by cp.async.bulk,
    // 4 / 5 = force lockstep / free-running.  NB2_VELOCITY_KERNEL overrides (A/B runs, tests).
    int velocity_kernel = 2;
    bool poison_rows = false;      // NB2_POISON_ROWS: fill the row planes with NaN bits before every assembly (tests)
    size_t smem_optin = 0;         // cudaDevAttrMaxSharedMemoryPerBlockOptin
    bool staged_attr = false, staged_pos_attr = false, bulk_attr = false, lockstep_attr = false;
    DevBuf<float4> p_hdr;          // [5][n_ghdr_max] coloured position groups: bodies + collider-to-body poses
    size_t n_ghdr_max = 0;
    void* host_hdr = nullptr;      // pinned copy of vs.hdr (launch-geometry hint, never waited for)
    DevBuf<unsigned int> bal;       // groups per colour while balancing
    bool pos_early_exit = true;     // NB2_POS_EARLY_EXIT=0: always run every position iteration (the exactness test)
    int ref_blocks = 0;             // NB2_REF_BLOCKS: cap on the blocks of the reference-order solve kernels (0 = automatic)
    bool kempe = true;              // NB2_KEMPE: Kempe-chain stage of the colouring (schedule.cu)
    DevBuf<int> col_edge;           // its per-body, per-colour group table
    bool producer_attr = false;     // k_generate_manifolds: dynamic shared memory size set
    void* mb = nullptr;             // MbState (multibody.cu): reduced-coordinate multibodies, SURVEY 8 f3
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
NB2_D float4 ldcg4(const float4* p) { return __ldcg(p); }
NB2_D void stcg4(float4* p, float4 v) { __stcg(p, v); }

// Software grid barrier for cooperatively launched kernels: a monotonically
// increasing arrival counter (zeroed before the launch).
// Regions of Context::barrier (words): the scheduling kernels' counter and flags, the velocity kernel's counter, the
// position kernel's counter and sweep flags.  The step zeroes the last two with one memset (api.cu).
#define NB2_BARRIER_WORDS 32
#define NB2_BARRIER_VELOCITY 8
#define NB2_BARRIER_POSITION 16
struct GridBarrier {
    unsigned int* counter;
    unsigned int target;
    __device__ void init(unsigned int* c) {
        counter = c;
        target = 0;
    }
    __device__ void sync() {
        __syncthreads();
        if (gridDim.x == 1) return;  // a single block: bar.sync already orders its threads' global accesses
        if (threadIdx.x == 0) {
            target += gridDim.x;
            // arrive with release semantics (orders this block's earlier writes, made visible to
            // thread 0 by the bar.sync above), then poll with acquire loads: no separate full
            // fences on either side of the atomic
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned int v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            } while (v < target);
        }
        __syncthreads();
    }
#ifdef NB2_TRACE
    // the same barrier, block 0 / thread 0 leaving timestamps: [0] whole block arrived, [1] arrival published (the
    // release fence has drained), [2] every block arrived
    __device__ void sync_traced(unsigned long long* ts) {
        __syncthreads();
        if (gridDim.x == 1) return;
        if (threadIdx.x == 0) {
            if (ts) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[0]));
            target += gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            if (ts) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[1]));
            unsigned int v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            } while (v < target);
            if (ts) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[2]));
        }
        __syncthreads();
    }
#endif
};

// Live body state as the kernels see it.
struct BodyPose {
    Pose pose;
    Vec3 com;
};
NB2_D Vec3 f4_xyz(float4 v) { return mk3(v.x, v.y, v.z); }
NB2_D float4 xyz_f4(Vec3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
NB2_D Quat f4_quat(float4 v) { return mkq(v.x, v.y, v.z, v.w); }
NB2_D float4 quat_f4(Quat q) { return make_float4(q.i, q.j, q.k, q.w); }

// RigidBody::apply_displacement (src/object/rigid_body.rs:371-381): the
// displacement (lin, ang) is applied about the centre of mass, then the com is
// refreshed from the new pose (set_position, :305-315).
NB2_D void apply_displacement(BodyPose* b, Vec3 local_com, Vec3 lin, Vec3 ang) {
    Quat dr = quat_from_scaled_axis(ang);
    Pose wrt_com;
    wrt_com.t = (b->com + lin) + quat_rotate(dr, -b->com);
    wrt_com.r = dr;
    b->pose = pose_mul(wrt_com, b->pose);
    b->com = pose_point(b->pose, local_com);
}

// ---------------------------------------------------------------------------
// launchers (one per .cu file group)
// ---------------------------------------------------------------------------
// bodies.cu
int launch_unpack_bodies(Context* ctx);
int launch_apply_effective_status(Context* ctx);
int launch_update_activation(Context* ctx, float mix, const int32_t* to_activate, uint32_t n_list);
int launch_label_islands(Context* ctx, int* d_labels, unsigned int* d_rows);
int launch_refresh_dynamics(Context* ctx);
int launch_integrate(Context* ctx, bool kinematic_only);
int launch_pack_states(Context* ctx, nb2_body_state* d_out, uint32_t first, uint32_t n);
int launch_unpack_states(Context* ctx, const nb2_body_state* d_in, uint32_t first, uint32_t n);
int launch_stats(Context* ctx, int mode);
int launch_body_stats(Context* ctx, double* d_energy, unsigned int* d_non_finite);
int launch_validate_inputs(Context* ctx);
int launch_validate_joints(Context* ctx);
// multibody.cu
int mb_upload(Context* ctx, const nb2_multibody* mbs, uint32_t n_mb, const nb2_mb_link* links, uint32_t n_links);
int mb_download_links(Context* ctx, nb2_mb_link* out, uint32_t n);
int mb_launch_refresh(Context* ctx);
int mb_launch_velocity(Context* ctx);
int mb_launch_position(Context* ctx);
int mb_count(Context* ctx);
void mb_release(Context* ctx);
void mb_invalidate(Context* ctx);
// schedule.cu
int exclusive_scan_u32(Context* ctx, const unsigned int* in, unsigned int* out, size_t n);
int launch_build_items(Context* ctx, int mode);
int launch_schedule(Context* ctx, Sched* s, int mode);
int query_coop_limits(Context* ctx);
// assemble.cu
int launch_assemble(Context* ctx, int mode);
int launch_cache_impulses(Context* ctx, int mode);
// narrowphase.cu
int launch_upload_colliders(Context* ctx, const nb2_collider* colliders, uint32_t n);
int launch_detect_pairs(Context* ctx, float prediction, float search, unsigned int flip_permille, uint32_t* out_pairs);
int launch_generate_manifolds(Context* ctx);
int launch_update_contacts(Context* ctx, const nb2_contact_update* updates, uint32_t n);
// solve.cu
int launch_velocity_solve(Context* ctx, int mode);
int launch_position_solve(Context* ctx, int mode);

}  // namespace nb2
