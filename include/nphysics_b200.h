/*
 * nphysics_b200.h -- C ABI of the B200-native MoreauJeanSolver hot path.
 *
 * This is the drop-in boundary.  A host (the Rust
 * `MechanicalWorld::step` wrapper shown in INTEGRATION.md, the C++ mirror in
 * nphysics_b200/host/, or the ctypes binding used by tests and bench.py) calls
 * these entry points in place of
 *
 *     self.solver.step(counters, bodies, colliders, constraints, manifolds,
 *                      island, island_joints, parameters, coefficients)
 *         -- reference: src/world/mechanical_world.rs:316-326,
 *                       src/solver/moreau_jean_solver.rs:47-90
 *
 * Conventions
 *   - every function returns an `int`: NB2_OK (0) or a negative nb2_error.
 *     Nothing panics/aborts; nb2_last_error() gives a message for the last
 *     failure on that context (or on the calling thread when ctx == NULL).
 *   - plain pointers + counts; caller-owned host buffers; context-owned device
 *     buffers; no hidden global state; one context per world / GPU.
 *   - f32 throughout (every reference example instantiates N = f32,
 *     examples3d/pyramid3.rs:90); 3D only (DIM = 3, SPATIAL_DIM = 6,
 *     src/lib.rs:264-346).
 *   - vectors are xyz; rotations are unit quaternions stored (i, j, k, w) like
 *     nalgebra's UnitQuaternion; an isometry is translation[3] then rotation[4];
 *     6-vectors are linear[3] then angular[3] (src/algebra/velocity3.rs:9-16).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     NB2_ERR_NO_DEVICE / NB2_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef NPHYSICS_B200_H
#define NPHYSICS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB2_ABI_VERSION 1

/* ------------------------------------------------------------------ errors */
typedef enum nb2_error {
    NB2_OK = 0,
    NB2_ERR_INVALID_ARGUMENT = -1,
    NB2_ERR_NO_DEVICE = -2,        /* no CUDA device / not sm_100         */
    NB2_ERR_CUDA = -3,             /* a CUDA runtime call failed          */
    NB2_ERR_OUT_OF_MEMORY = -4,
    NB2_ERR_BAD_INDEX = -5,        /* body index out of range in a record */
    NB2_ERR_UNSUPPORTED = -6,      /* e.g. a row between a body and itself */
    NB2_ERR_TOO_MANY_COLOURS = -7, /* colouring needs > NB2_MAX_COLOURS   */
    NB2_ERR_NOT_READY = -8,        /* step before bodies/params uploaded  */
    NB2_ERR_NON_FINITE = -9        /* NaN/Inf detected by nb2_compute_stats */
} nb2_error;

/* ------------------------------------------------------------------ params */
/* Mirrors IntegrationParameters (src/solver/integration_parameters.rs:5-76,
 * defaults :169-189) plus MechanicalWorld.gravity
 * (src/world/mechanical_world.rs:55-68).  inv_dt is derived exactly as
 * set_dt does (:141-153): 0 when dt == 0, else 1/dt. */
typedef struct nb2_params {
    float dt;
    float erp;
    float warmstart_coeff;
    float restitution_velocity_threshold;
    float allowed_linear_error;
    float allowed_angular_error;
    float max_linear_correction;
    float max_angular_correction;
    float max_stabilization_multiplier; /* kept for API parity; unused by the reference too */
    uint32_t max_velocity_iterations;
    uint32_t max_position_iterations;
    uint32_t max_ccd_position_iterations; /* CCD is out of scope: carried, ignored */
    uint32_t max_ccd_substeps;            /* carried, ignored */
    float gravity[3];
} nb2_params;

/* ------------------------------------------------------------------ bodies */
/* BodyStatus (src/object/body.rs:50-59). */
typedef enum nb2_body_status {
    NB2_BODY_DISABLED = 0,
    NB2_BODY_STATIC = 1,
    NB2_BODY_DYNAMIC = 2,
    NB2_BODY_KINEMATIC = 3,
    /* The record stands for one link of a multibody (nb2_mb_link.body points at it): manifolds and colliders refer
     * to the link through this body index; its pose, centre of mass and velocity are OUTPUTS, written by the
     * multibody kinematics (MultibodyLink::position / velocity, src/object/multibody_link.rs). */
    NB2_BODY_MULTIBODY_LINK = 4
} nb2_body_status;

#define NB2_BODY_FLAG_GRAVITY 1u /* RigidBody.gravity_enabled (rigid_body.rs:43) */

/* The fields of RigidBody the hot path reads (src/object/rigid_body.rs:26-50).
 * The Ground body (src/object/ground.rs) is a NB2_BODY_STATIC record.  Body
 * index = position in the uploaded array; it plays the role of the handle and
 * (x6) of the companion/assembly id (moreau_jean_solver.rs:143-152). */
typedef struct nb2_body {
    float position[7];             /* Isometry: translation xyz, quaternion ijkw */
    float velocity[6];             /* linear, angular */
    float local_com[3];
    float mass;                    /* local_inertia.linear */
    float local_inertia[9];        /* local_inertia.angular, row-major 3x3 */
    float external_forces[6];      /* Force: linear, torque */
    float linear_damping;
    float angular_damping;
    float max_linear_velocity;     /* FLT_MAX = unlimited (rigid_body.rs:72-73) */
    float max_angular_velocity;
    float jacobian_mask[6];        /* 1 = free dof, 0 = kinematic dof (rigid_body.rs:105-121) */
    uint32_t status;               /* nb2_body_status */
    uint32_t flags;                /* NB2_BODY_FLAG_* */
} nb2_body;

/* What a step changes in a body: pose and velocity. */
typedef struct nb2_body_state {
    float position[7];
    float velocity[6];
} nb2_body_state;

/* ActivationStatus of a body (src/object/body.rs:65-125): the sleeping state that
 * ActivationManager::update (src/detection/activation_manager.rs:60-201) maintains. */
typedef struct nb2_activation {
    float threshold; /* deactivation threshold; < 0 = None: the body never sleeps (default 0.01, body.rs:72-74) */
    float energy;    /* low-pass filtered squared generalized velocity; 0 = asleep (new_active: 4 * threshold) */
} nb2_activation;

/* --------------------------------------------------------------- manifolds */
/* LocalShapeApproximation geometry tags of ncollide's ContactKinematic
 * (SURVEY.md appendix B). */
typedef enum nb2_kinematic_geom {
    NB2_GEOM_POINT = 0,
    NB2_GEOM_LINE = 1,  /* Line/Line, Line/Point, Point/Line are resolved through their closest points
                         * (separated branch); Line/Plane, Plane/Line yield none */
    NB2_GEOM_PLANE = 2
} nb2_kinematic_geom;

/* One ColliderContactManifold (src/detection/collider_contact_manifold.rs:9-24)
 * with everything the solver reads from the two colliders pre-resolved:
 * body handles (:52-58), margins (src/object/collider.rs:165-167),
 * position_wrt_body_part (nonlinear_sor_prox.rs:191-197) and the per-pair
 * material combination (src/material/material.rs:134-177) -- use
 * nb2_combine_materials() to produce friction/restitution/surface_velocity. */
typedef struct nb2_manifold {
    int32_t body1;
    int32_t body2;
    uint32_t first_contact; /* index into the contact array */
    uint32_t num_contacts;
    float margin1;
    float margin2;
    float friction;
    float restitution;
    float surface_velocity[3]; /* props1.surface_velocity - props2.surface_velocity, world */
    float coll1_wrt_body[7];   /* isometry of collider 1 in body 1's frame */
    float coll2_wrt_body[7];
} nb2_manifold;

/* One TrackedContact: ncollide Contact {world1, world2, normal, depth}, its
 * ContactId (stand-in: any stable non-zero 64-bit key; 0 = null id, never
 * cached -- signorini_coulomb_pyramid_model.rs:233-260) and its
 * ContactKinematic (points/directions in the COLLIDER's local frame). */
typedef struct nb2_contact {
    float world1[3];
    float world2[3];
    float normal[3]; /* unit, from shape 1 toward shape 2 */
    float depth;     /* > 0 = penetration (before margins) */
    uint64_t key;
    float local1[3];
    float local2[3];
    float dir1[3]; /* plane normal / line direction of approx1 (unit), unused for points */
    float dir2[3];
    float dilation1; /* kinematic margin1 (before the collider margin is added) */
    float dilation2;
    uint8_t geom1; /* nb2_kinematic_geom */
    uint8_t geom2;
    uint8_t pad_[6];
} nb2_contact;

/* The per-step part of a TrackedContact: what ncollide recomputes for a contact that persists from
 * one step to the next (the leading 40 bytes of nb2_contact).  Everything else in the record -- id,
 * ContactKinematic local points / directions / dilations / geometry tags -- stays as uploaded. */
typedef struct nb2_contact_update {
    float world1[3];
    float world2[3];
    float normal[3];
    float depth;
} nb2_contact_update;

/* ----------------------------------------------------------------- colliders */
/* A cuboid collider for the device manifold producer: ncollide's
 * Cuboid::half_extents plus the Collider fields the contact path reads -- margin
 * (src/object/collider.rs:165-167, default 0.01: :457-479), position_wrt_body_part and the BasicMaterial
 * (src/material/basic_material.rs:14-43) with its combine modes (0 Average, 1 Min, 2 Multiply, 3 Max:
 * src/material/material.rs:72-86).  Surface velocities are not carried (BasicMaterial's default: none). */
typedef struct nb2_collider {
    float half_extents[3];
    float margin;
    float translation_wrt_body[3];
    float friction;
    float rotation_wrt_body[4]; /* unit quaternion ijkw */
    float restitution;
    int32_t body;               /* index of the body the collider is attached to */
    uint8_t friction_mode;
    uint8_t restitution_mode;
    uint8_t pad_[2];
    uint32_t flags;             /* reserved, 0 */
} nb2_collider;

/* ------------------------------------------------------------------ joints */
/* Constraint-based joints (src/joint/ *_constraint.rs). */
typedef enum nb2_joint_type {
    NB2_JOINT_BALL = 0,        /* ball_constraint.rs        */
    NB2_JOINT_REVOLUTE = 1,    /* revolute_constraint.rs    */
    NB2_JOINT_PRISMATIC = 2,   /* prismatic_constraint.rs   */
    NB2_JOINT_UNIVERSAL = 3,   /* universal_constraint.rs   */
    NB2_JOINT_PLANAR = 4,      /* planar_constraint.rs      */
    NB2_JOINT_RECTANGULAR = 5, /* rectangular_constraint.rs */
    NB2_JOINT_PIN_SLOT = 6,    /* pin_slot_constraint.rs    */
    NB2_JOINT_CYLINDRICAL = 7, /* cylindrical_constraint.rs */
    NB2_JOINT_FIXED = 8,       /* fixed_constraint.rs       */
    NB2_JOINT_CARTESIAN = 9,   /* cartesian_constraint.rs   */
    NB2_JOINT_TYPE_COUNT = 10
} nb2_joint_type;

#define NB2_JOINT_FLAG_MIN_OFFSET 1u /* prismatic min_offset is Some(..) */
#define NB2_JOINT_FLAG_MAX_OFFSET 2u /* prismatic max_offset is Some(..) */

typedef struct nb2_joint {
    uint32_t type; /* nb2_joint_type */
    int32_t body1;
    int32_t body2;
    uint32_t flags;
    float anchor1[3]; /* local to body 1 */
    float anchor2[3]; /* local to body 2 */
    float axis1[3];   /* unit, local to body 1: axis1 / axis_v1 (pin-slot) */
    float axis2[3];   /* unit, local to body 2: axis2 / axis_w2 (pin-slot) */
    float axis3[3];   /* pin-slot only: axis_w1, local to body 1 */
    float ref_frame1[4]; /* fixed/cartesian: quaternion ijkw */
    float ref_frame2[4];
    float angle;      /* universal */
    float min_offset; /* prismatic */
    float max_offset;
    float break_force_squared;  /* FLT_MAX = unbreakable */
    float break_torque_squared;
    /* cached impulses = the joint's warm start: lin[0..3], ang[3..6], limit[6].
     * Slot usage per type follows each reference cache_impulses verbatim. */
    float impulses[7];
    uint32_t broken; /* set by the solver when a break threshold is exceeded */
} nb2_joint;

/* -------------------------------------------------------------- multibodies */
/* Reduced-coordinate articulated bodies (src/object/multibody.rs, multibody_link.rs).
 * A multibody is a tree of links; link k hangs on its parent through a joint that owns ndofs generalized
 * coordinates.  Links of one multibody are contiguous and a parent precedes its children
 * (MultibodyDesc::build order, multibody.rs:1433-1470). */
typedef enum nb2_mb_joint_type {
    NB2_MBJ_FREE = 0,      /* free_joint.rs: 6 dofs (only as the root) */
    NB2_MBJ_BALL = 1,      /* ball_joint.rs: 3 dofs                     */
    NB2_MBJ_REVOLUTE = 2,  /* revolute_joint.rs: 1 dof, limits + motor  */
    NB2_MBJ_PRISMATIC = 3, /* prismatic_joint.rs: 1 dof, limits + motor */
    NB2_MBJ_FIXED = 4,     /* fixed_joint.rs: 0 dofs                    */
    NB2_MBJ_TYPE_COUNT = 5
} nb2_mb_joint_type;
#define NB2_MBJ_FLAG_MIN 1u   /* min_angle / min_offset is Some(..) */
#define NB2_MBJ_FLAG_MAX 2u   /* max_angle / max_offset is Some(..) */
#define NB2_MBJ_FLAG_MOTOR 4u /* JointMotor.enabled (joint_motor.rs) */
#define NB2_MB_MAX_DOFS 64    /* generalized coordinates of one multibody */

typedef struct nb2_mb_link {
    int32_t multibody;      /* index into the nb2_multibody array */
    int32_t parent;         /* link index WITHIN the multibody, -1 = root (its parent is the ground) */
    uint32_t joint_type;    /* nb2_mb_joint_type */
    uint32_t flags;         /* NB2_MBJ_FLAG_* */
    int32_t body;           /* the link's NB2_BODY_MULTIBODY_LINK record (mass properties are read from it:
                               mass, local_inertia, local_com) */
    float parent_shift[3];  /* MultibodyLink.parent_shift: joint anchor in the parent's frame */
    float body_shift[3];    /* MultibodyLink.body_shift: joint anchor to the link's origin, in the link's frame */
    float axis[3];          /* revolute / prismatic: unit axis */
    /* joint coordinates.  Free: translation xyz + quaternion ijkw; Ball: quaternion ijkw in [0..4);
     * Revolute: angle in [0]; Prismatic: offset in [0]; Fixed: body_to_parent isometry (t, q). */
    float coords[7];
    float velocity[6];      /* generalized velocities of the link's dofs (first ndofs entries) */
    float damping[6];       /* Multibody.damping of the link's dofs (Joint::default_damping: 0.1 for ball / revolute) */
    float min_pos, max_pos; /* unit joints: limits */
    float motor_velocity;   /* JointMotor.desired_velocity */
    float motor_max_velocity;
    float motor_max_force;
    float impulses[3];      /* unit joints: cached impulses of the motor, min and max rows (unit_joint.rs:77,113,153) */
} nb2_mb_link;

typedef struct nb2_multibody {
    uint32_t first_link;
    uint32_t n_links;
    uint32_t flags;         /* NB2_BODY_FLAG_GRAVITY */
    uint32_t reserved;
} nb2_multibody;

/* ------------------------------------------------------------------- step */
typedef enum nb2_step_mode {
    /* Replays the reference's sequential Gauss-Seidel order exactly (rows are
     * levelised into waves that preserve every body-wise dependency). */
    NB2_MODE_REFERENCE_ORDER = 0,
    /* Production: graph-coloured batches, colouring computed on device. */
    NB2_MODE_COLOURED = 1
} nb2_step_mode;

typedef struct nb2_stats {
    uint32_t n_bodies;
    uint32_t n_dynamic_bodies;
    uint32_t n_manifolds;
    uint32_t n_contacts;
    uint32_t n_joints;
    uint32_t n_rows_two_body; /* velocity rows between two dynamic bodies (R2) */
    uint32_t n_rows_ground;   /* velocity rows with one dynamic side (RG) */
    uint32_t n_phases_velocity; /* colours, or levels in reference-order mode */
    uint32_t n_phases_position;
    uint32_t n_broken_joints;
    uint32_t non_finite;      /* count of NaN/Inf body states seen */
    uint32_t schedule_verdict; /* coloured mode, last step: 0 cached schedule reused, 1 coloured from scratch,
                                * 2 cached + one refinement pass, 3 edited in place (a few groups changed) */
    float residual_max;       /* max |lambda - prox(lambda - r*(J dv + rhs))| over velocity rows */
    float residual_rms;
    float max_penetration;    /* max contact depth (incl. margins) at the final poses */
    float kinetic_energy;
    /* stage timers with the reference's Counters names (src/counters/mod.rs),
     * milliseconds, valid when timing was enabled for the step */
    float t_assembly_ms;
    float t_velocity_resolution_ms;
    float t_velocity_update_ms;
    float t_position_resolution_ms;
    float t_step_ms;
    float pad2_;
} nb2_stats;

typedef struct nb2_context nb2_context;

/* Basic (non-compute) queries: usable without a GPU. */
int nb2_abi_version(void);
const char* nb2_error_string(int err);
/* IntegrationParameters::default() + gravity (0,-9.81,0)
 * (integration_parameters.rs:169-189, examples3d/pyramid3.rs:21). */
int nb2_default_params(nb2_params* out);
/* sizeof() of each ABI struct, for binding self-checks:
 * which: 0 params, 1 body, 2 body_state, 3 manifold, 4 contact, 5 joint, 6 stats, 7 activation,
 * 8 contact_update, 9 collider, 10 mb_link, 11 multibody */
int nb2_sizeof(int which);
/* Material::combine for two BasicMaterials (material.rs:72-86,134-177,
 * basic_material.rs:30-43).  mode: 0 Average, 1 Min, 2 Multiply, 3 Max.
 * surface velocities are already rotated to world space (or NULL = none). */
int nb2_combine_materials(float friction1, int friction_mode1, float restitution1, int restitution_mode1,
                          const float* surface_velocity1, float friction2, int friction_mode2,
                          float restitution2, int restitution_mode2, const float* surface_velocity2,
                          float* out_friction, float* out_restitution, float* out_surface_velocity3);

/* Context lifetime.  `device` is a CUDA ordinal.  `stream` may be NULL (the
 * context creates its own) or an existing cudaStream_t on that device. */
int nb2_create(int device, void* stream, nb2_context** out_ctx);
int nb2_destroy(nb2_context* ctx);
const char* nb2_last_error(const nb2_context* ctx);

int nb2_set_params(nb2_context* ctx, const nb2_params* params);
int nb2_get_params(const nb2_context* ctx, nb2_params* out);
/* Enables CUDA-event stage timers (adds event records, no extra syncs until
 * nb2_get_stats). */
int nb2_enable_timers(nb2_context* ctx, int enabled);

/* The coloured schedule is cached across steps while the conflict graph (body pair, row count and
 * type of every constraint group) is unchanged; the check runs on device every step.  enabled = 0
 * forces a fresh colouring each step (default: enabled). */
int nb2_set_schedule_cache(nb2_context* ctx, int enabled);

/* Device representation of contact groups in coloured mode: 0 (default) = 100-byte rows
 * [J1|J2|M^-1 J1 . ang|M^-1 J2 . ang|header] streamed from HBM (the linear half of M^-1 J is rebuilt
 * from the inverse mass); 1 = compact 80-byte contact records from which the same rows are rebuilt
 * in registers (bit-identical rows, a quarter of the bytes, more ALU). */
int nb2_set_contact_layout(nb2_context* ctx, int layout);

/* Replace the whole body set (n >= 1).  Marks dynamics dirty, like
 * update_status = all() on a fresh body (rigid_body.rs:80). */
int nb2_upload_bodies(nb2_context* ctx, const nb2_body* bodies, uint32_t n);
/* Overwrite pose+velocity of bodies [first, first+n).  Asynchronous on the context's stream, like
 * nb2_upload_manifolds: a pinned host array must stay valid until the next nb2_synchronize / download. */
int nb2_upload_body_states(nb2_context* ctx, const nb2_body_state* states, uint32_t first, uint32_t n);
/* The contact set for the next step ("uploaded once per step"). */
int nb2_upload_manifolds(nb2_context* ctx, const nb2_manifold* manifolds, uint32_t n_manifolds,
                         const nb2_contact* contacts, uint32_t n_contacts);
/* Contacts that persist from the previous step (same manifold list, same contact order, same ids):
 * only the 40 bytes per contact that changed are uploaded and scattered into the device records of the
 * last nb2_upload_manifolds.  n_contacts must equal that upload's count; when a contact appears or
 * disappears the host uploads the whole set again.  Asynchronous like nb2_upload_manifolds. */
int nb2_update_contacts(nb2_context* ctx, const nb2_contact_update* updates, uint32_t n_contacts);

/* ---- Device manifold producer (replaces the ncollide side of
 * src/world/geometrical_world.rs:285-320 for cuboid piles).  With it a step needs no contact upload at
 * all: bodies, colliders and persistent feature pairs live on the device.
 *
 * nb2_upload_colliders: the cuboid colliders (any number per body; bodies must be uploaded first).
 * nb2_detect_pairs: broad phase + feature discovery at the CURRENT device poses: every (dynamic,
 *   dynamic) collider pair whose centres are within `search_radius` (< 0: (2 r_max + reach) * sqrt(2) *
 *   1.01, r_max the largest half extent of a dynamic collider) and every (non-dynamic, dynamic) pair,
 *   that face each other along one axis within reach = margin1 + margin2 + 2 * linear_prediction
 *   (collider.rs:542-549) with overlapping projections on the other two.  `flip_permille` of the pairs
 *   (a hash of the pair) list the Point side first (exercises Point/Plane kinematics).  Pairs are
 *   persistent until the next call; *out_pairs (may be NULL) receives their number.  Synchronises.
 * nb2_generate_manifolds: one manifold per pair at the current device poses -- the <= 4 corners of the
 *   faces' overlap rectangle closer than `reach`, contact i of pair p carrying the id 4 p + i + 1 -- written
 *   where nb2_upload_manifolds would have put them: manifold p owns the contact slots [4p, 4p + 4).
 *   Call it before every nb2_step instead of nb2_upload_manifolds.  Asynchronous.
 * nb2_download_manifolds: the device-side contact set of the next step (from either source), e.g. to
 *   feed a host-side consumer or the test oracle.  Writes min(capacity, n) records of each kind. */
int nb2_upload_colliders(nb2_context* ctx, const nb2_collider* colliders, uint32_t n_colliders);
int nb2_detect_pairs(nb2_context* ctx, float linear_prediction, float search_radius, uint32_t flip_permille,
                     uint32_t* out_pairs);
int nb2_generate_manifolds(nb2_context* ctx);
int nb2_download_manifolds(nb2_context* ctx, nb2_manifold* out_manifolds, uint32_t manifold_capacity,
                           nb2_contact* out_contacts, uint32_t contact_capacity, uint32_t* out_n_manifolds,
                           uint32_t* out_n_contacts);

/* The active joint set, in island_joints order.  Cached impulses/broken flags
 * in the records seed the device copy. */
int nb2_upload_joints(nb2_context* ctx, const nb2_joint* joints, uint32_t n_joints);
/* Multibodies (Multibody / MultibodyDesc::build, src/object/multibody.rs:1311-1470).  Upload after the bodies: every
 * link names its NB2_BODY_MULTIBODY_LINK record.  Only dynamic multibodies; they never sleep.  The step writes the
 * links' joint coordinates, generalized velocities and cached impulses back into the device copy
 * (nb2_download_multibody_links) and the links' world poses / velocities into their body records
 * (nb2_download_body_states). */
int nb2_upload_multibodies(nb2_context* ctx, const nb2_multibody* multibodies, uint32_t n_multibodies,
                           const nb2_mb_link* links, uint32_t n_links);
int nb2_download_multibody_links(nb2_context* ctx, nb2_mb_link* out, uint32_t n_links);
/* ContactModel selection (src/solver/contact_model.rs:13-37, MoreauJeanSolver::set_contact_model,
 * moreau_jean_solver.rs:42-44).  The reference ships two models:
 *   NB2_CONTACT_SIGNORINI_COULOMB_PYRAMID (default, moreau_jean_solver.rs:29-40): one unilateral normal
 *     row + two friction-pyramid rows per contact, every contact of a manifold makes rows
 *     (signorini_coulomb_pyramid_model.rs:56-224);
 *   NB2_CONTACT_SIGNORINI: frictionless -- the normal row and the position row only, and only for contacts
 *     with depth + margin1 + margin2 >= 0 (signorini_model.rs:141-150, 200-298); a contact that is
 *     inactive in a step keeps its cached impulse.
 * Changing the model drops the impulse cache (the cache belongs to the model object). */
typedef enum nb2_contact_model {
    NB2_CONTACT_SIGNORINI_COULOMB_PYRAMID = 0,
    NB2_CONTACT_SIGNORINI = 1
} nb2_contact_model;
int nb2_set_contact_model(nb2_context* ctx, int model);
/* Drop the contact impulse cache (a fresh ContactModel). */
int nb2_clear_impulse_cache(nb2_context* ctx);

/* Sleeping.  Off until nb2_upload_activation is called: every body is then awake for
 * good, as with set_deactivation_threshold(None).  n must equal the body count. */
int nb2_upload_activation(nb2_context* ctx, const nb2_activation* activation, uint32_t n);
/* ActivationManager::update (activation_manager.rs:60-201; call site mechanical_world.rs:265-272),
 * to be called between nb2_upload_manifolds / nb2_upload_joints and nb2_step: energy low-pass of the
 * awake dynamic bodies (mix_factor: 0.01 in the reference, mechanical_world.rs:80), deferred
 * activations (`to_activate`, body indices, may be NULL), islands = connected components over
 * dynamic and kinematic bodies joined by the uploaded manifolds that hold at least one contact and by
 * the unbroken joints, then an island whose bodies are all below their thresholds goes to sleep
 * (velocities zeroed, rigid_body.rs:396-400) and any other island is woken up.  Sleeping dynamic
 * bodies are left out of the following steps exactly like the reference leaves them out of
 * active_bodies and of the manifold list (mechanical_world.rs:287-300).  Asynchronous. */
int nb2_update_activation(nb2_context* ctx, float mix_factor, const int32_t* to_activate, uint32_t n_to_activate);
int nb2_download_activation(nb2_context* ctx, nb2_activation* out, uint32_t n);

/* Island labelling for sharding (the union-find of
 * src/detection/activation_manager.rs:122-159 restricted to dynamic bodies, :141-145): out_labels[i] =
 * the smallest body index of body i's connected component over the current manifolds (with or without
 * contacts: potential pairs keep their bodies together) and unbroken joints, -1 for a non-dynamic body;
 * out_rows[i] (may be NULL) = velocity rows of the constraint groups booked on body i (a group is booked
 * on its first dynamic body), i.e. the weights a host bin-packs islands onto GPUs with.  n must equal
 * the body count.  Synchronises. */
int nb2_label_islands(nb2_context* ctx, int32_t* out_labels, uint32_t* out_rows, uint32_t n);

/* One MoreauJeanSolver::step on the uploaded inputs, followed by the
 * kinematic-body integration and end-of-step dynamics refresh of
 * mechanical_world.rs:328-346.  Asynchronous on the context's stream. */
int nb2_step(nb2_context* ctx, int mode);
/* One MoreauJeanSolver::step_ccd (moreau_jean_solver.rs:94-127): a CCD sub-step on the uploaded inputs --
 * assemble, position resolution FIRST, then velocity resolution and integration; impulses are not cached
 * and kinematic bodies are not integrated.  The caller is the CCD driver (solve_ccd,
 * mechanical_world.rs:561-908: times of impact, frozen bodies, sub-step parameters with
 * warmstart_coeff = 0), which is not part of this library.  Asynchronous. */
int nb2_step_ccd(nb2_context* ctx, int mode);
int nb2_synchronize(nb2_context* ctx);

int nb2_download_body_states(nb2_context* ctx, nb2_body_state* out, uint32_t first, uint32_t n);
/* (lambda_n, lambda_t1, lambda_t2) per contact, in the order of the last upload. */
int nb2_download_contact_impulses(nb2_context* ctx, float* out3, uint32_t n_contacts);
/* Joint records with updated cached impulses and broken flags. */
int nb2_download_joints(nb2_context* ctx, nb2_joint* out, uint32_t n_joints);
/* Runs the residual / penetration / energy reduction for the last step and
 * returns it with the counters and timers. */
int nb2_get_stats(nb2_context* ctx, nb2_stats* out);
/* Stage timers of the last step only (needs nb2_enable_timers): waits for the step's last event
 * and writes {assembly, velocity_resolution, velocity_update, position_resolution, step,
 * velocity_kernel, position_kernel, schedule} in milliseconds.  No reductions are launched. */
int nb2_get_timers(nb2_context* ctx, float* out8);
/* The schedule of the last step, one entry per constraint group (the rows of <= 4 contacts of a
 * manifold, or the rows of one joint): its phase (colour in NB2_MODE_COLOURED, level in
 * NB2_MODE_REFERENCE_ORDER; -1 = group not scheduled: no dynamic side, broken joint) and the DYNAMIC
 * body of each side (-1 = static / kinematic / sleeping side: such sides never conflict).  Two groups
 * of one colour never share a dynamic body; tests assert exactly that.  Writes min(capacity, n) entries
 * and the group count to *out_n.  Any output array may be NULL. */
int nb2_download_schedule(nb2_context* ctx, int32_t* out_phase, int32_t* out_body1, int32_t* out_body2,
                          uint32_t capacity, uint32_t* out_n);
/* Number of kernels this context launched since creation (bench.py's
 * gpu_launches). */
int nb2_launch_count(const nb2_context* ctx, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* NPHYSICS_B200_H */
