// The extern "C" boundary (include/nphysics_b200.h) and the step orchestration.
//
// nb2_step replaces MoreauJeanSolver::step (src/solver/moreau_jean_solver.rs:47-90) together with
// the per-body work MechanicalWorld::step does around it (update_dynamics / update_acceleration
// before, kinematic integration after: src/world/mechanical_world.rs:230-243, 328-346).
#include <stdarg.h>
#include <stdlib.h>

#include <new>

#include "solver.cuh"

namespace nb2 {

static thread_local char g_thread_error[512] = "";

int set_error(Context* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) {
        strncpy(ctx->last_error, buf, sizeof(ctx->last_error) - 1);
        ctx->last_error[sizeof(ctx->last_error) - 1] = 0;
    }
    strncpy(g_thread_error, buf, sizeof(g_thread_error) - 1);
    g_thread_error[sizeof(g_thread_error) - 1] = 0;
    return code;
}

static void release_all(Context* c) {
    mb_release(c);
    c->raw.release(); c->pos.release(); c->pos_t.p.p = c->pos_q.p.p = nullptr; c->vel.release(); c->com_im.release();
    c->inv_i.release(); c->ext.release(); c->lam.release(); c->b_status.release(); c->joints.release();
    c->manifolds.release(); c->contacts.release(); c->c_manifold.release(); c->chunk_base.release();
    c->chunk_manifold.release(); c->col_edge.release();
    for (int k = 0; k < 2; ++k) {
        c->imp[k].release();
        c->ckey[k].release();
        c->ht_keys[k].release();
        c->ht_imps[k].release();
    }
    c->vs.release(); c->ps.release(); c->deg.release(); c->adj_off.release(); c->cursor.release();
    c->adj.release(); c->pred_a.release(); c->pred_b.release(); c->level.release(); c->adj_off_p.release();
    c->adj_p.release(); c->cmask.release(); c->best.release(); c->scan_tmp.release(); c->barrier.release();
    c->r_jac.release(); c->r_hdr.release(); c->r_imp.release(); c->p_row.release();
    c->stat_f.release(); c->stat_u.release(); c->flags.release(); c->stage_states.release();
    c->c_geo.release(); c->bal.release(); c->p_hdr.release();
    c->true_status.release(); c->isl_labels.release(); c->act.release(); c->cc_parent.release(); c->cc_can.release(); c->wake_list.release();
    c->colliders.release(); c->coll_world.release(); c->np_is_big.release(); c->np_big_off.release(); c->np_par.release();
    c->pair_cnt.release(); c->pair_off.release(); c->grid_count.release(); c->grid_off.release(); c->grid_cursor.release();
    c->np_big_list.release(); c->grid_entries.release(); c->pair_feat.release(); c->pair_q.release();
    c->contact_updates.release();
    if (c->host_hdr) cudaFreeHost(c->host_hdr);
    c->host_hdr = nullptr;
}

static int check_flags(Context* ctx) {
    // input-validation bits + schedule overflow, read after a synchronisation point
    unsigned int f = 0, fl[4] = {0, 0, 0, 0};
    if (ctx->flags.p) {
        NB2_CUDA(ctx, cudaMemcpyAsync(fl, ctx->flags.p, sizeof(fl), cudaMemcpyDeviceToHost, ctx->stream));
    }
    SchedHeader hv, hp;
    memset(&hv, 0, sizeof(hv));
    memset(&hp, 0, sizeof(hp));
    if (ctx->stepped && ctx->vs.hdr.p)
        NB2_CUDA(ctx, cudaMemcpyAsync(&hv, ctx->vs.hdr.p, sizeof(hv), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->stepped && ctx->last_mode == NB2_MODE_REFERENCE_ORDER && ctx->ps.hdr.p)
        NB2_CUDA(ctx, cudaMemcpyAsync(&hp, ctx->ps.hdr.p, sizeof(hp), cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    f = fl[0];
    ctx->last_stats.schedule_verdict = ctx->last_mode == NB2_MODE_COLOURED ? (fl[1] > 3u ? 1u : fl[1]) : 0u;
    ctx->last_stats.n_phases_velocity = hv.n_phases;
    ctx->last_stats.n_phases_position = ctx->last_mode == NB2_MODE_REFERENCE_ORDER ? hp.n_phases : hv.n_phases;
    if (f & 1u) return set_error(ctx, NB2_ERR_BAD_INDEX, "a manifold/joint record referenced a body or contact out of range");
    if (f & 2u) return set_error(ctx, NB2_ERR_UNSUPPORTED, "a manifold/joint connects a body to itself (unsupported)");
    if (f & 0x400u) return set_error(ctx, NB2_ERR_BAD_INDEX, "a multibody link names a body that is not a NB2_BODY_MULTIBODY_LINK record");
    if (f & 0x100u)
        return set_error(ctx, NB2_ERR_UNSUPPORTED, "a manifold couples a multibody with a dynamic rigid body: such rows are not solved (DESIGN.md section 8b)");
    if (f & 0x200u) return set_error(ctx, NB2_ERR_UNSUPPORTED, "a multibody touches more manifolds, or a component of multibodies holds more members / coordinates, than the multibody path holds (DESIGN.md section 8b)");
    if ((hv.overflow | hp.overflow) & 1u)
        return set_error(ctx, NB2_ERR_TOO_MANY_COLOURS, "colouring needs more than %d colours", NB2_MAX_COLOURS);
    if ((hv.overflow | hp.overflow) & 2u) return set_error(ctx, NB2_ERR_CUDA, "internal: phase index overflow");
    return NB2_OK;
}

// MoreauJeanSolver::step_ccd (moreau_jean_solver.rs:94-127): the same stages in the order assemble,
// position, velocity, integrate; no impulse caching, no kinematic integration (the CCD driver around it,
// mechanical_world.rs:561-908, owns those).
static int do_step_ccd(Context* ctx, int mode) {
    ctx->step_layout = (mode != NB2_MODE_REFERENCE_ORDER && ctx->contact_layout == 1 && ctx->contact_model == 0) ? 1 : 0;
    ctx->cur = 1 - ctx->cur;  // assembly warm-starts from the buffer "before cur": point it at the last one written
    NB2_CUDA(ctx, cudaMemsetAsync(ctx->barrier.p + NB2_BARRIER_VELOCITY, 0, 16 * sizeof(unsigned int), ctx->stream));
    NB2_TRY(launch_refresh_dynamics(ctx));
    NB2_TRY(mb_launch_refresh(ctx));  // multibodies: kinematics, dynamics, accelerations; their links' body records follow
    // chunks (groups of <= 4 contacts): at most one per manifold plus one per four contacts; the device producer's
    // manifolds own exactly one each
    ctx->max_chunks = ctx->manifolds_from_producer ? (size_t)ctx->n_manifolds : (size_t)ctx->n_manifolds + ctx->n_contacts / NB2_CHUNK;
    NB2_TRY(launch_build_items(ctx, mode));
    NB2_TRY(launch_schedule(ctx, &ctx->vs, mode));
    if (mode == NB2_MODE_REFERENCE_ORDER) NB2_TRY(launch_schedule(ctx, &ctx->ps, mode));
    NB2_TRY(launch_assemble(ctx, mode));
    NB2_TRY(launch_position_solve(ctx, mode));
    NB2_TRY(launch_velocity_solve(ctx, mode));
    NB2_TRY(launch_integrate(ctx, false));
    ctx->cur = 1 - ctx->cur;  // nothing was cached: the next step still reads the same buffer
    ctx->ev_valid = false;
    ctx->last_mode = mode;
    ctx->stepped = true;
    return NB2_OK;
}

static int do_step(Context* ctx, int mode) {
    const bool ref = mode == NB2_MODE_REFERENCE_ORDER;
    ctx->cur = 1 - ctx->cur;
    ctx->step_layout = (!ref && ctx->contact_layout == 1 && ctx->contact_model == 0) ? 1 : 0;
    const bool tm = ctx->timers;
    if (tm && !ctx->ev.created) {
        for (int k = 0; k < 12; ++k) NB2_CUDA(ctx, cudaEventCreate(&ctx->ev.e[k]));
        ctx->ev.created = true;
    }
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[0], ctx->stream));
    // the grid-barrier counters of the velocity and the position kernel (their own regions: one memset per step)
    NB2_CUDA(ctx, cudaMemsetAsync(ctx->barrier.p + NB2_BARRIER_VELOCITY, 0, 16 * sizeof(unsigned int), ctx->stream));
    // ---- dynamics refresh + assembly (Counters: "assembly")
    NB2_TRY(launch_refresh_dynamics(ctx));
    NB2_TRY(mb_launch_refresh(ctx));  // multibodies: kinematics, dynamics, accelerations; their links' body records follow
    // chunks (groups of <= 4 contacts): at most one per manifold plus one per four contacts; the device producer's
    // manifolds own exactly one each
    ctx->max_chunks = ctx->manifolds_from_producer ? (size_t)ctx->n_manifolds : (size_t)ctx->n_manifolds + ctx->n_contacts / NB2_CHUNK;
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[10], ctx->stream));
    NB2_TRY(launch_build_items(ctx, mode));
    NB2_TRY(launch_schedule(ctx, &ctx->vs, mode));
    if (ref) NB2_TRY(launch_schedule(ctx, &ctx->ps, mode));
    if (!ref) {
        // colour count of this step for the next step's launch geometry (a hint: never waited for)
        if (!ctx->host_hdr) {
            NB2_CUDA(ctx, cudaHostAlloc(&ctx->host_hdr, sizeof(SchedHeader), cudaHostAllocDefault));
            memset(ctx->host_hdr, 0, sizeof(SchedHeader));
        }
        NB2_CUDA(ctx, cudaMemcpyAsync(ctx->host_hdr, ctx->vs.hdr.p, sizeof(SchedHeader), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[11], ctx->stream));
    NB2_TRY(launch_assemble(ctx, mode));
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[1], ctx->stream));
    // ---- velocity resolution + impulse caching (Counters: "velocity resolution")
    NB2_TRY(launch_velocity_solve(ctx, mode));
    NB2_TRY(launch_cache_impulses(ctx, mode));
    NB2_TRY(mb_launch_velocity(ctx));  // multibodies: rows, velocity resolution, impulse caching, integration of the joints
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[2], ctx->stream));
    // ---- velocity update + integration (Counters: "velocity update")
    NB2_TRY(launch_integrate(ctx, false));
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[3], ctx->stream));
    // ---- position resolution
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[8], ctx->stream));
    NB2_TRY(launch_position_solve(ctx, mode));
    NB2_TRY(mb_launch_position(ctx));  // multibodies: position resolution, then the end-of-step kinematics
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[9], ctx->stream));
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[4], ctx->stream));
    // ---- kinematic bodies (mechanical_world.rs:328-332)
    if (ctx->n_kinematic) NB2_TRY(launch_integrate(ctx, true));
    if (tm) NB2_CUDA(ctx, cudaEventRecord(ctx->ev.e[5], ctx->stream));
    ctx->ev_valid = tm;
    ctx->last_mode = mode;
    ctx->stepped = true;
    return NB2_OK;
}

}  // namespace nb2

using namespace nb2;

#define NB2_CHECK_CTX(ctx)                                   \
    do {                                                     \
        if (!(ctx)) return set_error(nullptr, NB2_ERR_INVALID_ARGUMENT, "null context"); \
    } while (0)

struct nb2_context {
    Context c;
};

extern "C" {

int nb2_abi_version(void) { return NB2_ABI_VERSION; }

const char* nb2_error_string(int err) {
    switch (err) {
        case NB2_OK: return "ok";
        case NB2_ERR_INVALID_ARGUMENT: return "invalid argument";
        case NB2_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
        case NB2_ERR_CUDA: return "CUDA error";
        case NB2_ERR_OUT_OF_MEMORY: return "out of device memory";
        case NB2_ERR_BAD_INDEX: return "record index out of range";
        case NB2_ERR_UNSUPPORTED: return "unsupported input";
        case NB2_ERR_TOO_MANY_COLOURS: return "too many colours";
        case NB2_ERR_NOT_READY: return "bodies/params not uploaded";
        case NB2_ERR_NON_FINITE: return "non-finite body state";
        default: return "unknown error";
    }
}

int nb2_default_params(nb2_params* p) {
    if (!p) return NB2_ERR_INVALID_ARGUMENT;
    memset(p, 0, sizeof(*p));
    p->dt = 1.0f / 60.0f;
    p->erp = 0.2f;
    p->warmstart_coeff = 1.0f;
    p->restitution_velocity_threshold = 1.0f;
    p->allowed_linear_error = 0.001f;
    p->allowed_angular_error = 0.001f;
    p->max_linear_correction = 0.2f;
    p->max_angular_correction = 0.2f;
    p->max_stabilization_multiplier = 0.2f;
    p->max_velocity_iterations = 8;
    p->max_position_iterations = 3;
    p->max_ccd_position_iterations = 10;
    p->max_ccd_substeps = 1;
    p->gravity[0] = 0.f;
    p->gravity[1] = -9.81f;
    p->gravity[2] = 0.f;
    return NB2_OK;
}

int nb2_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(nb2_params);
        case 1: return (int)sizeof(nb2_body);
        case 2: return (int)sizeof(nb2_body_state);
        case 3: return (int)sizeof(nb2_manifold);
        case 4: return (int)sizeof(nb2_contact);
        case 5: return (int)sizeof(nb2_joint);
        case 6: return (int)sizeof(nb2_stats);
        case 7: return (int)sizeof(nb2_activation);
        case 8: return (int)sizeof(nb2_contact_update);
        case 9: return (int)sizeof(nb2_collider);
        case 10: return (int)sizeof(nb2_mb_link);
        case 11: return (int)sizeof(nb2_multibody);
        default: return NB2_ERR_INVALID_ARGUMENT;
    }
}

// MaterialCombineMode::combine (src/material/material.rs:72-86): precedence Max > Multiply > Min > Average.
static float combine_coeff(float a, int ma, float b, int mb) {
    if (ma == 3 || mb == 3) return a > b ? a : b;
    if (ma == 2 || mb == 2) return a * b;
    if (ma == 1 || mb == 1) return a < b ? a : b;
    return (a + b) * 0.5f;
}
int nb2_combine_materials(float friction1, int friction_mode1, float restitution1, int restitution_mode1,
                          const float* sv1, float friction2, int friction_mode2, float restitution2,
                          int restitution_mode2, const float* sv2, float* out_friction, float* out_restitution,
                          float* out_sv) {
    if (!out_friction || !out_restitution || !out_sv) return NB2_ERR_INVALID_ARGUMENT;
    *out_restitution = combine_coeff(restitution1, restitution_mode1, restitution2, restitution_mode2);
    *out_friction = combine_coeff(friction1, friction_mode1, friction2, friction_mode2);
    for (int k = 0; k < 3; ++k) out_sv[k] = (sv1 ? sv1[k] : 0.f) - (sv2 ? sv2[k] : 0.f);  // material.rs:174
    return NB2_OK;
}

const char* nb2_last_error(const nb2_context* ctx) { return ctx ? ctx->c.last_error : g_thread_error; }

int nb2_create(int device, void* stream, nb2_context** out) {
    if (!out) return set_error(nullptr, NB2_ERR_INVALID_ARGUMENT, "null out pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(nullptr, NB2_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return set_error(nullptr, NB2_ERR_INVALID_ARGUMENT, "bad device ordinal %d", device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return set_error(nullptr, NB2_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return set_error(nullptr, NB2_ERR_NO_DEVICE, "device %d is sm_%d%d; this build carries sm_100a code only", device,
                         prop.major, prop.minor);
    if (!prop.cooperativeLaunch) return set_error(nullptr, NB2_ERR_NO_DEVICE, "device lacks cooperative launch");
    nb2_context* h = new (std::nothrow) nb2_context();
    if (!h) return set_error(nullptr, NB2_ERR_OUT_OF_MEMORY, "host allocation failed");
    Context* ctx = &h->c;
    ctx->last_error[0] = 0;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    memset(&ctx->last_stats, 0, sizeof(ctx->last_stats));
    nb2_default_params(&ctx->params);
    ctx->inv_dt = 1.0f / ctx->params.dt;
    ctx->have_params = true;
    // developer knob for A/B measurements of the coloured velocity kernel variants (solver.cuh)
    if (const char* vk = getenv("NB2_VELOCITY_KERNEL")) ctx->velocity_kernel = atoi(vk);
    if (const char* pr = getenv("NB2_POISON_ROWS")) ctx->poison_rows = atoi(pr) != 0;
    if (const char* ps = getenv("NB2_POS_EARLY_EXIT")) ctx->pos_early_exit = atoi(ps) != 0;
    if (const char* rb = getenv("NB2_REF_BLOCKS")) ctx->ref_blocks = atoi(rb);
    if (const char* ic = getenv("NB2_INCREMENTAL_COLOURING")) ctx->incremental_colouring = atoi(ic) != 0;
    if (const char* kc = getenv("NB2_KEMPE")) ctx->kempe = atoi(kc) != 0;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete h;
        return set_error(nullptr, NB2_ERR_CUDA, "cudaSetDevice failed");
    }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
        ctx->own_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete h;
            return set_error(nullptr, NB2_ERR_CUDA, "cudaStreamCreate failed");
        }
        ctx->own_stream = true;
    }
    int rc = ctx->flags.reserve(ctx, 4);
    if (rc == NB2_OK) rc = ctx->barrier.reserve(ctx, NB2_BARRIER_WORDS);
    if (rc == NB2_OK && cudaMemsetAsync(ctx->flags.p, 0, 4 * sizeof(unsigned int), ctx->stream) != cudaSuccess)
        rc = NB2_ERR_CUDA;
    if (rc != NB2_OK) {
        release_all(ctx);
        delete h;
        return rc;
    }
    *out = h;
    return NB2_OK;
}

int nb2_destroy(nb2_context* h) {
    if (!h) return NB2_OK;
    Context* ctx = &h->c;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    release_all(ctx);
    if (ctx->ev.created)
        for (int k = 0; k < 12; ++k) cudaEventDestroy(ctx->ev.e[k]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete h;
    return NB2_OK;
}

int nb2_set_params(nb2_context* h, const nb2_params* p) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!p) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null params");
    if (!(p->dt >= 0.f)) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "The time-stepping length cannot be negative.");
    ctx->params = *p;
    ctx->inv_dt = p->dt == 0.f ? 0.f : 1.0f / p->dt;  // integration_parameters.rs:141-153
    ctx->have_params = true;
    return NB2_OK;
}

int nb2_get_params(const nb2_context* h, nb2_params* out) {
    if (!h || !out) return NB2_ERR_INVALID_ARGUMENT;
    *out = h->c.params;
    return NB2_OK;
}

int nb2_enable_timers(nb2_context* h, int enabled) {
    NB2_CHECK_CTX(h);
    h->c.timers = enabled != 0;
    return NB2_OK;
}

int nb2_set_schedule_cache(nb2_context* h, int enabled) {
    NB2_CHECK_CTX(h);
    h->c.schedule_cache = enabled != 0;
    return NB2_OK;
}

int nb2_set_contact_layout(nb2_context* h, int layout) {
    NB2_CHECK_CTX(h);
    if (layout != 0 && layout != 1) return set_error(&h->c, NB2_ERR_INVALID_ARGUMENT, "unknown contact layout %d", layout);
    h->c.contact_layout = layout;
    return NB2_OK;
}

int nb2_upload_bodies(nb2_context* h, const nb2_body* bodies, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!bodies || n == 0) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "empty body set");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // buffers may be reallocated
    NB2_TRY(ctx->raw.reserve(ctx, n));
    NB2_TRY(ctx->pos.reserve(ctx, 2 * (size_t)n));
    ctx->pos_t.p.p = ctx->pos.p;
    ctx->pos_q.p.p = ctx->pos.p + 1;
    NB2_TRY(ctx->vel.reserve(ctx, 2 * (size_t)n));
    NB2_TRY(ctx->com_im.reserve(ctx, n));
    NB2_TRY(ctx->inv_i.reserve(ctx, 3 * (size_t)n));
    NB2_TRY(ctx->ext.reserve(ctx, 2 * (size_t)n));
    NB2_TRY(ctx->lam.reserve(ctx, 2 * (size_t)n));
    NB2_TRY(ctx->b_status.reserve(ctx, n));
    NB2_TRY(ctx->true_status.reserve(ctx, n));
    if (n != ctx->n_bodies) ctx->sleeping = false;  // activation records of another body set
    ctx->n_bodies = n;
    uint32_t nd = 0, nk = 0;
    bool any_mask = false;
    for (uint32_t i = 0; i < n; ++i) {
        nd += bodies[i].status == NB2_BODY_DYNAMIC ? 1u : 0u;
        nk += bodies[i].status == NB2_BODY_KINEMATIC ? 1u : 0u;
        for (int k = 0; k < 6; ++k) any_mask |= bodies[i].jacobian_mask[k] != 1.0f;
    }
    ctx->n_dynamic = nd;
    ctx->n_kinematic = nk;
    ctx->any_mask = any_mask;
    NB2_CUDA(ctx, cudaMemcpyAsync(ctx->raw.p, bodies, (size_t)n * sizeof(nb2_body), cudaMemcpyHostToDevice, ctx->stream));
    NB2_TRY(launch_unpack_bodies(ctx));
    NB2_TRY(launch_apply_effective_status(ctx));  // sleeping bodies stay asleep across a re-upload of the same set
    // joints / manifolds referring to the old set are dropped
    ctx->n_manifolds = ctx->n_contacts = 0;
    ctx->n_joints = 0;
    mb_invalidate(ctx);  // multibody links name body records of the previous set
    ctx->n_colliders = ctx->n_pairs = 0;
    ctx->pairs_valid = false;
    ctx->stepped = false;
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller may free `bodies` on return
    return NB2_OK;
}

int nb2_upload_body_states(nb2_context* h, const nb2_body_state* states, uint32_t first, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!states && n) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null states");
    if ((uint64_t)first + n > ctx->n_bodies) return set_error(ctx, NB2_ERR_BAD_INDEX, "body range out of bounds");
    if (!n) return NB2_OK;
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((size_t)n > ctx->stage_states.cap) NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // about to reallocate
    NB2_TRY(ctx->stage_states.reserve(ctx, n));
    NB2_CUDA(ctx, cudaMemcpyAsync(ctx->stage_states.p, states, (size_t)n * sizeof(nb2_body_state),
                                  cudaMemcpyHostToDevice, ctx->stream));
    NB2_TRY(launch_unpack_states(ctx, ctx->stage_states.p, first, n));
    // asynchronous on the context's stream (pageable memory is staged by the runtime before the call returns;
    // a pinned array must stay valid until the next nb2_synchronize / download, like nb2_upload_manifolds')
    return NB2_OK;
}

int nb2_upload_manifolds(nb2_context* h, const nb2_manifold* manifolds, uint32_t nm, const nb2_contact* contacts,
                         uint32_t nc) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if ((nm && !manifolds) || (nc && !contacts)) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null manifold/contact array");
    if (!ctx->n_bodies) return set_error(ctx, NB2_ERR_NOT_READY, "upload bodies first");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((size_t)nm > ctx->manifolds.cap || (size_t)nc > ctx->contacts.cap)
        NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // about to reallocate
    NB2_TRY(ctx->manifolds.reserve(ctx, nm));
    NB2_TRY(ctx->contacts.reserve(ctx, nc));
    ctx->n_manifolds = nm;
    ctx->n_contacts = nc;
    ctx->manifolds_from_producer = false;
    if (nm)
        NB2_CUDA(ctx, cudaMemcpyAsync(ctx->manifolds.p, manifolds, (size_t)nm * sizeof(nb2_manifold),
                                      cudaMemcpyHostToDevice, ctx->stream));
    if (nc)
        NB2_CUDA(ctx, cudaMemcpyAsync(ctx->contacts.p, contacts, (size_t)nc * sizeof(nb2_contact),
                                      cudaMemcpyHostToDevice, ctx->stream));
    NB2_TRY(launch_validate_inputs(ctx));
    // asynchronous: the caller's arrays must stay valid until the next nb2_synchronize / download
    return NB2_OK;
}

int nb2_update_contacts(nb2_context* h, const nb2_contact_update* updates, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (n && !updates) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null contact updates");
    if (n != ctx->n_contacts)
        return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "contact updates: expected %u (the last uploaded contact set), got %u",
                         ctx->n_contacts, n);
    if (!n) return NB2_OK;
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((size_t)n > ctx->contact_updates.cap) NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return launch_update_contacts(ctx, updates, n);
}

int nb2_upload_colliders(nb2_context* h, const nb2_collider* colliders, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (n && !colliders) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null colliders");
    if (!ctx->n_bodies) return set_error(ctx, NB2_ERR_NOT_READY, "upload bodies before their colliders");
    for (uint32_t i = 0; i < n; ++i)
        if (colliders[i].body < 0 || (uint32_t)colliders[i].body >= ctx->n_bodies)
            return set_error(ctx, NB2_ERR_BAD_INDEX, "collider %u is attached to body %d (of %u)", i, colliders[i].body, ctx->n_bodies);
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return launch_upload_colliders(ctx, colliders, n);
}

int nb2_detect_pairs(nb2_context* h, float linear_prediction, float search_radius, uint32_t flip_permille, uint32_t* out_pairs) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (out_pairs) *out_pairs = 0;
    if (!ctx->n_colliders) return set_error(ctx, NB2_ERR_NOT_READY, "upload colliders first");
    if (!(linear_prediction >= 0.f)) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "the prediction distance must be >= 0");
    if (flip_permille > 1000u) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "flip_permille is a fraction of 1000");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_TRY(launch_detect_pairs(ctx, linear_prediction, search_radius, flip_permille, out_pairs));
    unsigned int f = 0;
    NB2_CUDA(ctx, cudaMemcpyAsync(&f, ctx->flags.p, sizeof(f), cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f & 4u) {
        unsigned int keep = f & ~4u;
        NB2_CUDA(ctx, cudaMemcpyAsync(ctx->flags.p, &keep, sizeof(keep), cudaMemcpyHostToDevice, ctx->stream));
        NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->pairs_valid = false;
        return set_error(ctx, NB2_ERR_UNSUPPORTED, "a collider has more contact partners than the producer tracks");
    }
    return NB2_OK;
}

int nb2_generate_manifolds(nb2_context* h) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!ctx->pairs_valid) return set_error(ctx, NB2_ERR_NOT_READY, "call nb2_detect_pairs first");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_generate_manifolds(ctx);
}

int nb2_download_manifolds(nb2_context* h, nb2_manifold* out_m, uint32_t cap_m, nb2_contact* out_c, uint32_t cap_c,
                           uint32_t* out_nm, uint32_t* out_nc) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (out_nm) *out_nm = ctx->n_manifolds;
    if (out_nc) *out_nc = ctx->n_contacts;
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t km = out_m ? (ctx->n_manifolds < cap_m ? ctx->n_manifolds : cap_m) : 0;
    const size_t kc = out_c ? (ctx->n_contacts < cap_c ? ctx->n_contacts : cap_c) : 0;
    if (km) NB2_CUDA(ctx, cudaMemcpyAsync(out_m, ctx->manifolds.p, km * sizeof(nb2_manifold), cudaMemcpyDeviceToHost, ctx->stream));
    if (kc) NB2_CUDA(ctx, cudaMemcpyAsync(out_c, ctx->contacts.p, kc * sizeof(nb2_contact), cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB2_OK;
}

int nb2_upload_joints(nb2_context* h, const nb2_joint* joints, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (n && !joints) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null joints");
    if (!ctx->n_bodies) return set_error(ctx, NB2_ERR_NOT_READY, "upload bodies first");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NB2_TRY(ctx->joints.reserve(ctx, n));
    ctx->n_joints = n;
    if (n) {
        NB2_CUDA(ctx, cudaMemcpyAsync(ctx->joints.p, joints, (size_t)n * sizeof(nb2_joint), cudaMemcpyHostToDevice,
                                      ctx->stream));
        NB2_TRY(launch_validate_joints(ctx));
        NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return NB2_OK;
}

int nb2_upload_multibodies(nb2_context* h, const nb2_multibody* multibodies, uint32_t n_multibodies, const nb2_mb_link* links,
                           uint32_t n_links) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    return mb_upload(ctx, multibodies, n_multibodies, links, n_links);
}

int nb2_download_multibody_links(nb2_context* h, nb2_mb_link* out, uint32_t n_links) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (n_links && !out) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null output");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    return mb_download_links(ctx, out, n_links);
}

int nb2_set_contact_model(nb2_context* h, int model) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (model != NB2_CONTACT_SIGNORINI_COULOMB_PYRAMID && model != NB2_CONTACT_SIGNORINI)
        return set_error(ctx, NB2_ERR_UNSUPPORTED, "unknown contact model %d (the reference ships two: 0 Signorini-Coulomb pyramid, 1 Signorini)", model);
    if (model != ctx->contact_model) {
        ctx->ht_cap[0] = ctx->ht_cap[1] = 0;
        ctx->imp_n[0] = ctx->imp_n[1] = 0;
        ctx->vs.cache_valid = false;  // the groups' row counts change
    }
    ctx->contact_model = model;
    return NB2_OK;
}

int nb2_clear_impulse_cache(nb2_context* h) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    ctx->ht_cap[0] = ctx->ht_cap[1] = 0;
    ctx->imp_n[0] = ctx->imp_n[1] = 0;
    return NB2_OK;
}

int nb2_upload_activation(nb2_context* h, const nb2_activation* activation, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!ctx->n_bodies) return set_error(ctx, NB2_ERR_NOT_READY, "upload bodies before their activation records");
    if (!activation || n != ctx->n_bodies)
        return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "activation records: expected %u, got %u", ctx->n_bodies, n);
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_TRY(ctx->act.reserve(ctx, n));
    static_assert(sizeof(nb2_activation) == sizeof(float2), "nb2_activation is copied as float2");
    NB2_CUDA(ctx, cudaMemcpyAsync(ctx->act.p, activation, (size_t)n * sizeof(nb2_activation), cudaMemcpyHostToDevice,
                                  ctx->stream));
    ctx->sleeping = true;
    NB2_TRY(launch_apply_effective_status(ctx));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller may free `activation` on return
    return NB2_OK;
}

int nb2_update_activation(nb2_context* h, float mix_factor, const int32_t* to_activate, uint32_t n_to_activate) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!ctx->sleeping) return set_error(ctx, NB2_ERR_NOT_READY, "sleeping is off: call nb2_upload_activation first");
    if (!(mix_factor >= 0.f)) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "the energy mixing factor must be >= 0");
    if (n_to_activate && !to_activate) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null activation list");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_update_activation(ctx, mix_factor, to_activate, n_to_activate);
}

int nb2_download_activation(nb2_context* h, nb2_activation* out, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!ctx->sleeping) return set_error(ctx, NB2_ERR_NOT_READY, "sleeping is off: call nb2_upload_activation first");
    if (!out || n > ctx->n_bodies) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "bad activation download range");
    if (!n) return NB2_OK;
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_CUDA(ctx, cudaMemcpyAsync(out, ctx->act.p, (size_t)n * sizeof(nb2_activation), cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB2_OK;
}

int nb2_label_islands(nb2_context* h, int32_t* out_labels, uint32_t* out_rows, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!ctx->n_bodies) return set_error(ctx, NB2_ERR_NOT_READY, "upload bodies first");
    if (n != ctx->n_bodies || !out_labels)
        return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "island labels: expected room for %u bodies, got %u", ctx->n_bodies, n);
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_TRY(ctx->isl_labels.reserve(ctx, n));
    NB2_TRY(ctx->cc_can.reserve(ctx, (size_t)n + 1));
    NB2_TRY(launch_label_islands(ctx, ctx->isl_labels.p, ctx->cc_can.p));
    NB2_CUDA(ctx, cudaMemcpyAsync(out_labels, ctx->isl_labels.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_rows)
        NB2_CUDA(ctx, cudaMemcpyAsync(out_rows, ctx->cc_can.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB2_OK;
}

int nb2_step(nb2_context* h, int mode) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (mode != NB2_MODE_REFERENCE_ORDER && mode != NB2_MODE_COLOURED)
        return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "unknown step mode %d", mode);
    if (!ctx->n_bodies || !ctx->have_params) return set_error(ctx, NB2_ERR_NOT_READY, "upload bodies and params before stepping");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    return do_step(ctx, mode);
}

int nb2_step_ccd(nb2_context* h, int mode) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (mode != NB2_MODE_REFERENCE_ORDER && mode != NB2_MODE_COLOURED)
        return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "unknown step mode %d", mode);
    if (!ctx->n_bodies || !ctx->have_params) return set_error(ctx, NB2_ERR_NOT_READY, "upload bodies and params before stepping");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    return do_step_ccd(ctx, mode);
}

int nb2_synchronize(nb2_context* h) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    return check_flags(ctx);
}

int nb2_download_body_states(nb2_context* h, nb2_body_state* out, uint32_t first, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!out && n) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null output");
    if ((uint64_t)first + n > ctx->n_bodies) return set_error(ctx, NB2_ERR_BAD_INDEX, "body range out of bounds");
    if (!n) return NB2_OK;
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((size_t)n > ctx->stage_states.cap) NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NB2_TRY(ctx->stage_states.reserve(ctx, n));
    NB2_TRY(launch_pack_states(ctx, ctx->stage_states.p, first, n));
    NB2_CUDA(ctx, cudaMemcpyAsync(out, ctx->stage_states.p, (size_t)n * sizeof(nb2_body_state), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB2_OK;
}

int nb2_download_contact_impulses(nb2_context* h, float* out3, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!n) return NB2_OK;
    if (!out3) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null output");
    if (!ctx->stepped || n > ctx->imp_n[ctx->cur]) return set_error(ctx, NB2_ERR_BAD_INDEX, "more impulses requested than contacts solved");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    // device rows are float4 (n, t1, t2, pad): copy with a 2D pitch conversion
    NB2_CUDA(ctx, cudaMemcpy2DAsync(out3, 3 * sizeof(float), ctx->imp[ctx->cur].p, sizeof(float4), 3 * sizeof(float), n,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB2_OK;
}

int nb2_download_joints(nb2_context* h, nb2_joint* out, uint32_t n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!n) return NB2_OK;
    if (!out) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null output");
    if (n > ctx->n_joints) return set_error(ctx, NB2_ERR_BAD_INDEX, "more joints requested than uploaded");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_CUDA(ctx, cudaMemcpyAsync(out, ctx->joints.p, (size_t)n * sizeof(nb2_joint), cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB2_OK;
}

int nb2_get_stats(nb2_context* h, nb2_stats* out) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!out) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null output");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    nb2_stats& s = ctx->last_stats;
    memset(&s, 0, sizeof(s));
    s.n_bodies = ctx->n_bodies;
    s.n_dynamic_bodies = ctx->n_dynamic;
    s.n_manifolds = ctx->n_manifolds;
    s.n_contacts = ctx->n_contacts;
    s.n_joints = ctx->n_joints;
    if (ctx->stepped) {
        NB2_TRY(launch_stats(ctx, ctx->last_mode));
        float f[16];
        unsigned int u[16];
        NB2_CUDA(ctx, cudaMemcpyAsync(f, ctx->stat_f.p, sizeof(f), cudaMemcpyDeviceToHost, ctx->stream));
        NB2_CUDA(ctx, cudaMemcpyAsync(u, ctx->stat_u.p, sizeof(u), cudaMemcpyDeviceToHost, ctx->stream));
        NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        double d[2];
        memcpy(d, f + 4, sizeof(d));
        s.residual_max = f[0];
        s.residual_rms = u[0] ? (float)sqrt(d[0] / (double)u[0]) : 0.f;
        {
            int key = (int)u[4];
            if (key == (int)0x80000000) {
                s.max_penetration = 0.f;
            } else {
                int bits = key >= 0 ? key : (key ^ 0x7FFFFFFF);
                memcpy(&s.max_penetration, &bits, sizeof(float));
            }
        }
        s.kinetic_energy = (float)d[1];
        s.n_rows_two_body = u[1];
        s.n_rows_ground = u[2];
        s.non_finite = u[3];
        s.n_broken_joints = u[5];
        if (ctx->ev_valid) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev.e[0], ctx->ev.e[1]);
            s.t_assembly_ms = ms;
            cudaEventElapsedTime(&ms, ctx->ev.e[1], ctx->ev.e[2]);
            s.t_velocity_resolution_ms = ms;
            cudaEventElapsedTime(&ms, ctx->ev.e[2], ctx->ev.e[3]);
            s.t_velocity_update_ms = ms;
            cudaEventElapsedTime(&ms, ctx->ev.e[3], ctx->ev.e[4]);
            s.t_position_resolution_ms = ms;
            cudaEventElapsedTime(&ms, ctx->ev.e[0], ctx->ev.e[5]);
            s.t_step_ms = ms;
        }
    }
    int rc = check_flags(ctx);
    *out = ctx->last_stats;
    if (rc != NB2_OK) return rc;
    if (out->non_finite) return set_error(ctx, NB2_ERR_NON_FINITE, "%u bodies have a non-finite state", out->non_finite);
    return NB2_OK;
}

int nb2_get_timers(nb2_context* h, float* out8) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!out8) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null output");
    if (!ctx->ev_valid) return set_error(ctx, NB2_ERR_NOT_READY, "timers were not enabled for the last step");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    NB2_CUDA(ctx, cudaEventSynchronize(ctx->ev.e[5]));
    const int pairs[8][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {0, 5}, {6, 7}, {8, 9}, {10, 11}};
    for (int k = 0; k < 8; ++k) NB2_CUDA(ctx, cudaEventElapsedTime(&out8[k], ctx->ev.e[pairs[k][0]], ctx->ev.e[pairs[k][1]]));
    return NB2_OK;
}

int nb2_download_schedule(nb2_context* h, int32_t* out_phase, int32_t* out_body1, int32_t* out_body2, uint32_t capacity,
                          uint32_t* out_n) {
    NB2_CHECK_CTX(h);
    Context* ctx = &h->c;
    if (!out_n) return set_error(ctx, NB2_ERR_INVALID_ARGUMENT, "null out_n");
    if (!ctx->stepped) return set_error(ctx, NB2_ERR_NOT_READY, "no step has been scheduled yet");
    NB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const Sched& s = ctx->vs;
    const size_t n = s.n_items;
    *out_n = (uint32_t)n;
    const size_t k = n < capacity ? n : capacity;
    if (k) {
        static_assert(sizeof(int) == sizeof(int32_t), "schedule arrays are copied as int32");
        if (out_phase) NB2_CUDA(ctx, cudaMemcpyAsync(out_phase, s.it_phase.p, k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_body1) NB2_CUDA(ctx, cudaMemcpyAsync(out_body1, s.it_a.p, k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_body2) NB2_CUDA(ctx, cudaMemcpyAsync(out_body2, s.it_b.p, k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        // unscheduled groups keep whatever phase an earlier step left: report them as -1
        int* ty = (int*)malloc(k * sizeof(int));
        if (!ty) return set_error(ctx, NB2_ERR_OUT_OF_MEMORY, "host allocation failed");
        cudaError_t e = cudaMemcpyAsync(ty, s.it_type.p, k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess && out_phase)
            for (size_t i = 0; i < k; ++i)
                if (ty[i] == NB2_ITEM_INVALID) out_phase[i] = -1;
        free(ty);
        NB2_CUDA(ctx, e);
    }
    return NB2_OK;
}

int nb2_launch_count(const nb2_context* h, uint64_t* out) {
    if (!h || !out) return NB2_ERR_INVALID_ARGUMENT;
    *out = h->c.launches;
    return NB2_OK;
}

}  // extern "C"
