// nphysics_b200.hpp -- C++ host mirror of the reference's interface for the solver hot path.
//
// The reference is Rust; this environment has no Rust toolchain, so the compiled host side above
// the C ABI is C++ (task rule 2).  Names, argument meaning and error behaviour follow the
// reference:
//   IntegrationParameters      src/solver/integration_parameters.rs:5-190
//   BodyStatus                 src/object/body.rs:50-59
//   RigidBodyDesc / RigidBody  src/object/rigid_body.rs:26-50, 949-1086
//   Ground                     src/object/ground.rs:15-40
//   DefaultBodySet             src/object/body_set.rs:61-175
//   *Constraint (joints)       src/joint/*_constraint.rs
//   FreeJoint / BallJoint / RevoluteJoint / PrismaticJoint / FixedJoint, MultibodyDesc, Multibody
//                              src/joint/*_joint.rs, src/object/multibody.rs:1311-1470
//   ColliderContactManifold    src/detection/collider_contact_manifold.rs:9-115
//   ContactModel / SignoriniCoulombPyramidModel   src/solver/contact_model.rs:13-37,
//                              src/solver/signorini_coulomb_pyramid_model.rs:19-54
//   MoreauJeanSolver           src/solver/moreau_jean_solver.rs:14-90
//   MechanicalWorld::step      src/world/mechanical_world.rs:182-396 (solver part only)
//   Counters                   src/counters/mod.rs:19-219 (solver stages)
// Everything funnels into include/nphysics_b200.h; there is no CPU path.
#pragma once
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nphysics_b200.h"

namespace nphysics {

using Vector3 = std::array<float, 3>;
using Quaternion = std::array<float, 4>;  // i, j, k, w (nalgebra storage order)

struct Isometry3 {
    Vector3 translation{0.f, 0.f, 0.f};
    Quaternion rotation{0.f, 0.f, 0.f, 1.f};
};

// The reference panics on invariant violations (assert!/unwrap); the mirror throws.
struct SolverError : std::runtime_error {
    int code;
    SolverError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// ---------------------------------------------------------------------------------------------
// integration_parameters.rs:5-190
class IntegrationParameters {
    float dt_ = 1.0f / 60.0f;
    float inv_dt_ = 60.0f;

  public:
    float erp = 0.2f;
    float warmstart_coeff = 1.0f;
    float restitution_velocity_threshold = 1.0f;
    float allowed_linear_error = 0.001f;
    float allowed_angular_error = 0.001f;
    float max_linear_correction = 0.2f;
    float max_angular_correction = 0.2f;
    float max_stabilization_multiplier = 0.2f;
    size_t max_velocity_iterations = 8;
    size_t max_position_iterations = 3;
    size_t max_ccd_position_iterations = 10;
    size_t max_ccd_substeps = 1;
    bool return_after_ccd_substep = false;
    bool multiple_ccd_substep_sensor_events_enabled = false;
    bool ccd_on_penetration_enabled = false;

    float dt() const { return dt_; }
    float inv_dt() const { return inv_dt_; }
    void set_dt(float dt) {  // :141-153
        if (!(dt >= 0.f)) throw SolverError(NB2_ERR_INVALID_ARGUMENT, "The time-stepping length cannot be negative.");
        dt_ = dt;
        inv_dt_ = dt == 0.f ? 0.f : 1.0f / dt;
    }
    void set_inv_dt(float inv_dt) {  // :156-166
        inv_dt_ = inv_dt;
        dt_ = inv_dt == 0.f ? 0.f : 1.0f / inv_dt;
    }
    nb2_params to_abi(const Vector3& gravity) const {
        nb2_params p;
        std::memset(&p, 0, sizeof(p));
        p.dt = dt_;
        p.erp = erp;
        p.warmstart_coeff = warmstart_coeff;
        p.restitution_velocity_threshold = restitution_velocity_threshold;
        p.allowed_linear_error = allowed_linear_error;
        p.allowed_angular_error = allowed_angular_error;
        p.max_linear_correction = max_linear_correction;
        p.max_angular_correction = max_angular_correction;
        p.max_stabilization_multiplier = max_stabilization_multiplier;
        p.max_velocity_iterations = (uint32_t)max_velocity_iterations;
        p.max_position_iterations = (uint32_t)max_position_iterations;
        p.max_ccd_position_iterations = (uint32_t)max_ccd_position_iterations;
        p.max_ccd_substeps = (uint32_t)max_ccd_substeps;
        for (int k = 0; k < 3; ++k) p.gravity[k] = gravity[k];
        return p;
    }
};

// ---------------------------------------------------------------------------------------------
enum class BodyStatus : uint32_t { Disabled = 0, Static = 1, Dynamic = 2, Kinematic = 3, MultibodyLink = 4 };  // body.rs:50-59 (+ the link proxy)

using DefaultBodyHandle = size_t;
struct BodyPartHandle {
    DefaultBodyHandle body;
    size_t part;
};

class RigidBody {
    friend class RigidBodyDesc;
    friend class DefaultBodySet;
    friend class MoreauJeanSolver;
    nb2_body rec_;
    nb2_activation act_;  // ActivationStatus (body.rs:65-125)

  public:
    RigidBody() {
        act_.threshold = 0.01f;  // ActivationStatus::new_active (body.rs:77-82)
        act_.energy = 0.04f;
        std::memset(&rec_, 0, sizeof(rec_));
        rec_.position[6] = 1.f;
        rec_.max_linear_velocity = FLT_MAX;
        rec_.max_angular_velocity = FLT_MAX;
        for (int k = 0; k < 6; ++k) rec_.jacobian_mask[k] = 1.f;
        rec_.status = NB2_BODY_DYNAMIC;
        rec_.flags = NB2_BODY_FLAG_GRAVITY;
    }
    Isometry3 position() const {
        Isometry3 p;
        for (int k = 0; k < 3; ++k) p.translation[k] = rec_.position[k];
        for (int k = 0; k < 4; ++k) p.rotation[k] = rec_.position[3 + k];
        return p;
    }
    void set_position(const Isometry3& p) {
        for (int k = 0; k < 3; ++k) rec_.position[k] = p.translation[k];
        for (int k = 0; k < 4; ++k) rec_.position[3 + k] = p.rotation[k];
    }
    Vector3 linear_velocity() const { return {rec_.velocity[0], rec_.velocity[1], rec_.velocity[2]}; }
    Vector3 angular_velocity() const { return {rec_.velocity[3], rec_.velocity[4], rec_.velocity[5]}; }
    void set_linear_velocity(const Vector3& v) { for (int k = 0; k < 3; ++k) rec_.velocity[k] = v[k]; }
    void set_angular_velocity(const Vector3& v) { for (int k = 0; k < 3; ++k) rec_.velocity[3 + k] = v[k]; }
    BodyStatus status() const { return (BodyStatus)rec_.status; }
    void set_status(BodyStatus s) { rec_.status = (uint32_t)s; }
    bool is_dynamic() const { return rec_.status == NB2_BODY_DYNAMIC; }
    size_t status_dependent_ndofs() const { return is_dynamic() ? 6 : 0; }  // body.rs:287-293
    float mass() const { return rec_.mass; }
    // ActivationStatus (body.rs:65-125, rigid_body.rs:391-405): a negative threshold stands for None
    bool is_active() const { return act_.energy != 0.f; }
    float activation_energy() const { return act_.energy; }
    void set_deactivation_threshold(float threshold_or_negative_for_none) { act_.threshold = threshold_or_negative_for_none; }
    float deactivation_threshold() const { return act_.threshold; }
    const nb2_body& record() const { return rec_; }
    nb2_body& record_mut() { return rec_; }
};

// rigid_body.rs:949-1086
class RigidBodyDesc {
    RigidBody rb_;

  public:
    RigidBodyDesc& translation(const Vector3& t) { for (int k = 0; k < 3; ++k) rb_.rec_.position[k] = t[k]; return *this; }
    RigidBodyDesc& rotation(const Quaternion& q) { for (int k = 0; k < 4; ++k) rb_.rec_.position[3 + k] = q[k]; return *this; }
    RigidBodyDesc& velocity(const Vector3& lin, const Vector3& ang) {
        for (int k = 0; k < 3; ++k) { rb_.rec_.velocity[k] = lin[k]; rb_.rec_.velocity[3 + k] = ang[k]; }
        return *this;
    }
    RigidBodyDesc& mass(float m) { rb_.rec_.mass = m; return *this; }
    RigidBodyDesc& angular_inertia(const std::array<float, 9>& row_major) {
        for (int k = 0; k < 9; ++k) rb_.rec_.local_inertia[k] = row_major[k];
        return *this;
    }
    RigidBodyDesc& local_center_of_mass(const Vector3& c) { for (int k = 0; k < 3; ++k) rb_.rec_.local_com[k] = c[k]; return *this; }
    RigidBodyDesc& linear_damping(float d) { rb_.rec_.linear_damping = d; return *this; }
    RigidBodyDesc& angular_damping(float d) { rb_.rec_.angular_damping = d; return *this; }
    RigidBodyDesc& max_linear_velocity(float v) { rb_.rec_.max_linear_velocity = v; return *this; }
    RigidBodyDesc& max_angular_velocity(float v) { rb_.rec_.max_angular_velocity = v; return *this; }
    RigidBodyDesc& gravity_enabled(bool on) {
        rb_.rec_.flags = on ? (rb_.rec_.flags | NB2_BODY_FLAG_GRAVITY) : (rb_.rec_.flags & ~NB2_BODY_FLAG_GRAVITY);
        return *this;
    }
    RigidBodyDesc& status(BodyStatus s) { rb_.rec_.status = (uint32_t)s; return *this; }
    /// rigid_body.rs desc `sleep_threshold`: negative = None (the body never sleeps).
    RigidBodyDesc& sleep_threshold(float threshold_or_negative_for_none) {
        rb_.act_.threshold = threshold_or_negative_for_none;
        return *this;
    }
    RigidBodyDesc& kinematic_translations(bool x, bool y, bool z) {
        rb_.rec_.jacobian_mask[0] = x ? 0.f : 1.f; rb_.rec_.jacobian_mask[1] = y ? 0.f : 1.f; rb_.rec_.jacobian_mask[2] = z ? 0.f : 1.f;
        return *this;
    }
    RigidBodyDesc& kinematic_rotations(bool x, bool y, bool z) {
        rb_.rec_.jacobian_mask[3] = x ? 0.f : 1.f; rb_.rec_.jacobian_mask[4] = y ? 0.f : 1.f; rb_.rec_.jacobian_mask[5] = z ? 0.f : 1.f;
        return *this;
    }
    /// Mass and angular inertia of a cuboid collider of the given density (volumetric_cuboid.rs:10-77).
    RigidBodyDesc& cuboid_collider(const Vector3& half_extents, float density) {
        const float hx = half_extents[0], hy = half_extents[1], hz = half_extents[2];
        const float m = density * 8.f * hx * hy * hz;
        rb_.rec_.mass = m;
        std::memset(rb_.rec_.local_inertia, 0, sizeof(rb_.rec_.local_inertia));
        rb_.rec_.local_inertia[0] = m * (4.f * hy * hy + 4.f * hz * hz) / 12.f;
        rb_.rec_.local_inertia[4] = m * (4.f * hx * hx + 4.f * hz * hz) / 12.f;
        rb_.rec_.local_inertia[8] = m * (4.f * hx * hx + 4.f * hy * hy) / 12.f;
        return *this;
    }
    RigidBody build() const { return rb_; }
};

// ground.rs: a body with no degree of freedom.
struct Ground {
    static RigidBody make() {
        RigidBody g = RigidBodyDesc().status(BodyStatus::Static).gravity_enabled(false).build();
        return g;
    }
};

// ---------------------------------------------------------------------------------------------
// Reduced-coordinate joints (src/joint/*_joint.rs): what a MultibodyDesc link hangs on.
struct Joint {
    nb2_mb_link rec;
    explicit Joint(uint32_t type) {
        std::memset(&rec, 0, sizeof(rec));
        rec.joint_type = type;
        rec.parent = -1;
        rec.body = -1;
        rec.axis[0] = 1.f;
        rec.motor_max_velocity = FLT_MAX;  // JointMotor::new (joint_motor.rs:17-24)
        rec.motor_max_force = FLT_MAX;
    }
};
struct FreeJoint : Joint {  // free_joint.rs:16-20
    explicit FreeJoint(const Isometry3& position) : Joint(NB2_MBJ_FREE) {
        for (int k = 0; k < 3; ++k) rec.coords[k] = position.translation[k];
        for (int k = 0; k < 4; ++k) rec.coords[3 + k] = position.rotation[k];
    }
};
struct BallJoint : Joint {  // ball_joint.rs:22-29 (identity rotation; default_damping 0.1, :89-91)
    BallJoint() : Joint(NB2_MBJ_BALL) {
        rec.coords[3] = 1.f;
        for (int k = 0; k < 3; ++k) rec.damping[k] = 0.1f;
    }
};
struct UnitJointBase : Joint {  // unit_joint.rs: limits and motor of a one-dof joint
    explicit UnitJointBase(uint32_t type) : Joint(type) {}
    void enable_min(float v) { rec.flags |= NB2_MBJ_FLAG_MIN; rec.min_pos = v; }
    void enable_max(float v) { rec.flags |= NB2_MBJ_FLAG_MAX; rec.max_pos = v; }
    void enable_motor() { rec.flags |= NB2_MBJ_FLAG_MOTOR; }
    void set_desired_motor_velocity(float v) { rec.motor_velocity = v; }
    void set_max_motor_force(float f) { rec.motor_max_force = f; }
};
struct RevoluteJoint : UnitJointBase {  // revolute_joint.rs:47-60 (default_damping 0.1, :214-216)
    RevoluteJoint(const Vector3& axis, float angle) : UnitJointBase(NB2_MBJ_REVOLUTE) {
        for (int k = 0; k < 3; ++k) rec.axis[k] = axis[k];
        rec.coords[0] = angle;
        rec.damping[0] = 0.1f;
    }
    void enable_min_angle(float v) { enable_min(v); }
    void enable_max_angle(float v) { enable_max(v); }
    void enable_angular_motor() { enable_motor(); }
    void set_desired_angular_motor_velocity(float v) { set_desired_motor_velocity(v); }
    void set_max_angular_motor_torque(float t) { set_max_motor_force(t); }
};
struct PrismaticJoint : UnitJointBase {  // prismatic_joint.rs:36-46
    PrismaticJoint(const Vector3& axis, float offset) : UnitJointBase(NB2_MBJ_PRISMATIC) {
        for (int k = 0; k < 3; ++k) rec.axis[k] = axis[k];
        rec.coords[0] = offset;
    }
    void enable_min_offset(float v) { enable_min(v); }
    void enable_max_offset(float v) { enable_max(v); }
    void enable_linear_motor() { enable_motor(); }
    void set_desired_linear_motor_velocity(float v) { set_desired_motor_velocity(v); }
    void set_max_linear_motor_force(float f) { set_max_motor_force(f); }
};
struct FixedJoint : Joint {  // fixed_joint.rs:14-19: the argument is the joint's pose in the body's frame; identity here
    FixedJoint() : Joint(NB2_MBJ_FIXED) { rec.coords[6] = 1.f; }
};

// MultibodyDesc (multibody.rs:1311-1470): a tree of links, each with its joint, shifts and mass properties.
class MultibodyDesc {
    Joint joint_;
    Vector3 parent_shift_{0.f, 0.f, 0.f}, body_shift_{0.f, 0.f, 0.f}, local_com_{0.f, 0.f, 0.f};
    float mass_ = 0.f;
    std::array<float, 9> inertia_{};
    std::vector<MultibodyDesc> children_;
    friend class Multibody;

  public:
    explicit MultibodyDesc(const Joint& joint) : joint_(joint) {}
    MultibodyDesc& add_child(const Joint& joint) {
        children_.emplace_back(joint);
        return children_.back();
    }
    MultibodyDesc& set_joint(const Joint& j) { joint_ = j; return *this; }
    MultibodyDesc& set_parent_shift(const Vector3& v) { parent_shift_ = v; return *this; }
    MultibodyDesc& set_body_shift(const Vector3& v) { body_shift_ = v; return *this; }
    /// what the colliders contribute in the reference (Body::add_local_inertia_and_com, multibody.rs:1154-1172)
    MultibodyDesc& local_mass_properties(float mass, const Vector3& local_com, const std::array<float, 9>& inertia_row_major) {
        mass_ = mass;
        local_com_ = local_com;
        inertia_ = inertia_row_major;
        return *this;
    }
    /// mass properties of a cuboid collider (volumetric_cuboid.rs:47-77)
    MultibodyDesc& cuboid(const Vector3& half_extents, float density) {
        const float hx = half_extents[0], hy = half_extents[1], hz = half_extents[2];
        mass_ = density * 8.f * hx * hy * hz;
        inertia_ = {};
        inertia_[0] = mass_ * (4.f * hy * hy + 4.f * hz * hz) / 12.f;
        inertia_[4] = mass_ * (4.f * hx * hx + 4.f * hz * hz) / 12.f;
        inertia_[8] = mass_ * (4.f * hx * hx + 4.f * hy * hy) / 12.f;
        return *this;
    }
};

// A built Multibody: its links in MultibodyDesc::build order (a parent before its children).
class Multibody {
    std::vector<nb2_mb_link> links_;
    std::vector<nb2_body> parts_;  // the NB2_BODY_MULTIBODY_LINK record of every link
    bool gravity_enabled_ = true;
    friend class DefaultBodySet;
    friend class MoreauJeanSolver;
    void add(const MultibodyDesc& d, int parent) {
        nb2_mb_link l = d.joint_.rec;
        l.parent = parent;
        for (int k = 0; k < 3; ++k) { l.parent_shift[k] = d.parent_shift_[k]; l.body_shift[k] = d.body_shift_[k]; }
        nb2_body b;
        std::memset(&b, 0, sizeof(b));
        b.position[6] = 1.f;
        b.status = NB2_BODY_MULTIBODY_LINK;
        b.mass = d.mass_;
        for (int k = 0; k < 3; ++k) b.local_com[k] = d.local_com_[k];
        for (int k = 0; k < 9; ++k) b.local_inertia[k] = d.inertia_[k];
        for (int k = 0; k < 6; ++k) b.jacobian_mask[k] = 1.f;
        b.max_linear_velocity = b.max_angular_velocity = FLT_MAX;
        const int me = (int)links_.size();
        links_.push_back(l);
        parts_.push_back(b);
        for (const MultibodyDesc& c : d.children_) add(c, me);
    }

  public:
    explicit Multibody(const MultibodyDesc& desc) { add(desc, -1); }
    size_t num_links() const { return links_.size(); }
    const nb2_mb_link& link(size_t i) const { return links_[i]; }
    /// the link's pose in the world (MultibodyLink::position), as of the last step
    Isometry3 link_position(size_t i) const {
        Isometry3 p;
        for (int k = 0; k < 3; ++k) p.translation[k] = parts_[i].position[k];
        for (int k = 0; k < 4; ++k) p.rotation[k] = parts_[i].position[3 + k];
        return p;
    }
    void enable_gravity(bool e) { gravity_enabled_ = e; }
};

// body_set.rs:61-175 (arena -> dense vector; the handle is the index)
class DefaultBodySet {
    std::vector<RigidBody> bodies_;
    std::vector<Multibody> multibodies_;
    std::vector<size_t> mb_first_body_;  // body index of link 0 of every multibody (BodyPartHandle(handle, i) = that + i)
    bool dirty_ = true;
    friend class MoreauJeanSolver;

  public:
    /// A Multibody takes one body slot per link; the returned handle is that of link 0 and
    /// BodyPartHandle{handle + i, 0} names link i for colliders and manifolds.
    DefaultBodyHandle insert_multibody(const Multibody& mb) {
        const size_t first = bodies_.size();
        for (size_t i = 0; i < mb.num_links(); ++i) {
            RigidBody proxy;
            proxy.rec_ = mb.parts_[i];
            proxy.act_.threshold = -1.f;  // multibodies never sleep
            bodies_.push_back(proxy);
        }
        multibodies_.push_back(mb);
        mb_first_body_.push_back(first);
        dirty_ = true;
        return first;
    }
    size_t num_multibodies() const { return multibodies_.size(); }
    const Multibody& multibody(size_t i) const { return multibodies_[i]; }

    DefaultBodyHandle insert(const RigidBody& b) {
        bodies_.push_back(b);
        dirty_ = true;
        return bodies_.size() - 1;
    }
    size_t len() const { return bodies_.size(); }
    const RigidBody* get(DefaultBodyHandle h) const { return h < bodies_.size() ? &bodies_[h] : nullptr; }
    RigidBody* get_mut(DefaultBodyHandle h) {
        if (h >= bodies_.size()) return nullptr;
        dirty_ = true;  // BodyUpdateStatus dirty bits (body.rs:449-527), coarse
        return &bodies_[h];
    }
};

// ---------------------------------------------------------------------------------------------
// joints: every *_constraint.rs constructor, producing the flat record
class JointConstraint {
  protected:
    nb2_joint rec_;
    JointConstraint(uint32_t type, BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& anchor2) {
        std::memset(&rec_, 0, sizeof(rec_));
        rec_.type = type;
        rec_.body1 = (int32_t)b1.body;
        rec_.body2 = (int32_t)b2.body;
        for (int k = 0; k < 3; ++k) { rec_.anchor1[k] = anchor1[k]; rec_.anchor2[k] = anchor2[k]; }
        rec_.axis1[0] = rec_.axis2[0] = rec_.axis3[0] = 1.f;
        rec_.ref_frame1[3] = rec_.ref_frame2[3] = 1.f;
        rec_.break_force_squared = FLT_MAX;
        rec_.break_torque_squared = FLT_MAX;
    }
    static void set3(float* d, const Vector3& v) { for (int k = 0; k < 3; ++k) d[k] = v[k]; }

  public:
    void set_break_force(float f) { rec_.break_force_squared = f * f; }
    void set_break_torque(float t) { rec_.break_torque_squared = t * t; }
    bool is_broken() const { return rec_.broken != 0; }
    const nb2_joint& record() const { return rec_; }
    nb2_joint& record_mut() { return rec_; }
};
struct BallConstraint : JointConstraint {  // ball_constraint.rs:27-45
    BallConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& anchor2)
        : JointConstraint(NB2_JOINT_BALL, b1, b2, anchor1, anchor2) {}
};
struct RevoluteConstraint : JointConstraint {  // revolute_constraint.rs:54-83
    RevoluteConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& axis1,
                       const Vector3& anchor2, const Vector3& axis2)
        : JointConstraint(NB2_JOINT_REVOLUTE, b1, b2, anchor1, anchor2) { set3(rec_.axis1, axis1); set3(rec_.axis2, axis2); }
};
struct PrismaticConstraint : JointConstraint {  // prismatic_constraint.rs:40-75
    PrismaticConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& axis1, const Vector3& anchor2)
        : JointConstraint(NB2_JOINT_PRISMATIC, b1, b2, anchor1, anchor2) { set3(rec_.axis1, axis1); }
    void enable_min_offset(float v) { rec_.flags |= NB2_JOINT_FLAG_MIN_OFFSET; rec_.min_offset = v; }
    void enable_max_offset(float v) { rec_.flags |= NB2_JOINT_FLAG_MAX_OFFSET; rec_.max_offset = v; }
    void disable_min_offset() { rec_.flags &= ~NB2_JOINT_FLAG_MIN_OFFSET; }
    void disable_max_offset() { rec_.flags &= ~NB2_JOINT_FLAG_MAX_OFFSET; }
};
struct UniversalConstraint : JointConstraint {  // universal_constraint.rs:30-60
    UniversalConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& axis1,
                        const Vector3& anchor2, const Vector3& axis2, float angle)
        : JointConstraint(NB2_JOINT_UNIVERSAL, b1, b2, anchor1, anchor2) { set3(rec_.axis1, axis1); set3(rec_.axis2, axis2); rec_.angle = angle; }
};
struct PlanarConstraint : JointConstraint {  // planar_constraint.rs:30-58
    PlanarConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& axis1,
                     const Vector3& anchor2, const Vector3& axis2)
        : JointConstraint(NB2_JOINT_PLANAR, b1, b2, anchor1, anchor2) { set3(rec_.axis1, axis1); set3(rec_.axis2, axis2); }
};
struct RectangularConstraint : JointConstraint {  // rectangular_constraint.rs:28-55
    RectangularConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& axis1, const Vector3& anchor2)
        : JointConstraint(NB2_JOINT_RECTANGULAR, b1, b2, anchor1, anchor2) { set3(rec_.axis1, axis1); }
};
struct PinSlotConstraint : JointConstraint {  // pin_slot_constraint.rs:35-75
    PinSlotConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& axis_v1,
                      const Vector3& axis_w1, const Vector3& anchor2, const Vector3& axis_w2)
        : JointConstraint(NB2_JOINT_PIN_SLOT, b1, b2, anchor1, anchor2) { set3(rec_.axis1, axis_v1); set3(rec_.axis3, axis_w1); set3(rec_.axis2, axis_w2); }
};
struct CylindricalConstraint : JointConstraint {  // cylindrical_constraint.rs:33-65
    CylindricalConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Vector3& axis1,
                          const Vector3& anchor2, const Vector3& axis2)
        : JointConstraint(NB2_JOINT_CYLINDRICAL, b1, b2, anchor1, anchor2) { set3(rec_.axis1, axis1); set3(rec_.axis2, axis2); }
};
struct FixedConstraint : JointConstraint {  // fixed_constraint.rs:30-60
    FixedConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Quaternion& ref_frame1,
                    const Vector3& anchor2, const Quaternion& ref_frame2)
        : JointConstraint(NB2_JOINT_FIXED, b1, b2, anchor1, anchor2) {
        for (int k = 0; k < 4; ++k) { rec_.ref_frame1[k] = ref_frame1[k]; rec_.ref_frame2[k] = ref_frame2[k]; }
    }
};
struct CartesianConstraint : JointConstraint {  // cartesian_constraint.rs:28-55
    CartesianConstraint(BodyPartHandle b1, BodyPartHandle b2, const Vector3& anchor1, const Quaternion& ref_frame1,
                        const Vector3& anchor2, const Quaternion& ref_frame2)
        : JointConstraint(NB2_JOINT_CARTESIAN, b1, b2, anchor1, anchor2) {
        for (int k = 0; k < 4; ++k) { rec_.ref_frame1[k] = ref_frame1[k]; rec_.ref_frame2[k] = ref_frame2[k]; }
    }
};

using DefaultJointConstraintHandle = size_t;
class DefaultJointConstraintSet {  // joint_constraint.rs:58-206
    std::vector<JointConstraint> joints_;
    bool dirty_ = true;
    friend class MoreauJeanSolver;

  public:
    DefaultJointConstraintHandle insert(const JointConstraint& j) {
        joints_.push_back(j);
        dirty_ = true;
        return joints_.size() - 1;
    }
    size_t len() const { return joints_.size(); }
    const JointConstraint* get(DefaultJointConstraintHandle h) const { return h < joints_.size() ? &joints_[h] : nullptr; }
};

// ---------------------------------------------------------------------------------------------
// Colliders and the geometrical world (src/object/collider.rs:457-560 ColliderDesc, src/object/collider_set.rs,
// src/world/geometrical_world.rs).  Collision detection is ncollide's in the reference; what runs here is the
// device manifold producer of SURVEY 8 f2 -- cuboid colliders, face-face contacts -- so that
// `mechanical_world.step(&mut geometrical_world, &mut bodies, &mut colliders, &mut joint_constraints, ...)`
// (mechanical_world.rs:182-188) has a counterpart with the same shape.  Other shapes: produce the manifolds with
// ncollide and call the overload of MechanicalWorld::step that takes them.
struct BasicMaterial {  // basic_material.rs:14-43 (defaults :46-50: restitution 0, friction 0.5, Average)
    float restitution = 0.f, friction = 0.5f;
    uint8_t restitution_combine_mode = 0, friction_combine_mode = 0;  // 0 Average, 1 Min, 2 Multiply, 3 Max
};
class Collider {
    friend class ColliderDesc;
    friend class DefaultColliderSet;
    friend class MoreauJeanSolver;
    nb2_collider rec_;
    float linear_prediction_ = 0.001f;  // collider.rs:457-479

  public:
    Collider() { std::memset(&rec_, 0, sizeof(rec_)); rec_.rotation_wrt_body[3] = 1.f; }
    DefaultBodyHandle body() const { return (DefaultBodyHandle)rec_.body; }
    float margin() const { return rec_.margin; }
    Vector3 half_extents() const { return {rec_.half_extents[0], rec_.half_extents[1], rec_.half_extents[2]}; }
};
class ColliderDesc {
    Collider c_;

  public:
    /// ColliderDesc::new(ShapeHandle::new(Cuboid::new(half_extents)))
    explicit ColliderDesc(const Vector3& cuboid_half_extents) {
        for (int k = 0; k < 3; ++k) c_.rec_.half_extents[k] = cuboid_half_extents[k];
        c_.rec_.margin = 0.01f;  // collider.rs:457-479
        const BasicMaterial m;
        material(m);
    }
    ColliderDesc& margin(float m) { c_.rec_.margin = m; return *this; }
    ColliderDesc& linear_prediction(float p) { c_.linear_prediction_ = p; return *this; }
    ColliderDesc& translation(const Vector3& t) {  // position of the collider in its body part's frame
        for (int k = 0; k < 3; ++k) c_.rec_.translation_wrt_body[k] = t[k];
        return *this;
    }
    ColliderDesc& material(const BasicMaterial& m) {
        c_.rec_.friction = m.friction;
        c_.rec_.restitution = m.restitution;
        c_.rec_.friction_mode = m.friction_combine_mode;
        c_.rec_.restitution_mode = m.restitution_combine_mode;
        return *this;
    }
    /// ColliderDesc::build(BodyPartHandle): link i of a multibody is BodyPartHandle{handle + i, 0}
    Collider build(const BodyPartHandle& part) const {
        Collider c = c_;
        c.rec_.body = (int32_t)(part.body + part.part);
        return c;
    }
};
using DefaultColliderHandle = size_t;
class DefaultColliderSet {
    std::vector<Collider> colliders_;
    bool dirty_ = true;
    friend class MoreauJeanSolver;

  public:
    DefaultColliderHandle insert(const Collider& c) {
        colliders_.push_back(c);
        dirty_ = true;
        return colliders_.size() - 1;
    }
    size_t len() const { return colliders_.size(); }
    const Collider* get(DefaultColliderHandle h) const { return h < colliders_.size() ? &colliders_[h] : nullptr; }
};
/// geometrical_world.rs: the broad phase (here: the persistent pairs of nb2_detect_pairs, found again every
/// `broad_phase_interval` steps from the poses of that moment) and the narrow phase (nb2_generate_manifolds, every step).
struct GeometricalWorld {
    uint32_t broad_phase_interval = 1;
    float search_radius = -1.f;  // < 0: derived from the largest dynamic collider
    uint32_t n_pairs = 0;        // pairs of the last broad phase
    uint64_t steps_ = 0;
};

// ---------------------------------------------------------------------------------------------
// The contact input (collider_contact_manifold.rs:9-24): one manifold and its tracked contacts.
struct ColliderContactManifold {
    nb2_manifold manifold;
    std::vector<nb2_contact> contacts;
};

// contact_model.rs:13-37.  Only the pyramid model exists on device.
struct ContactModel {
    virtual ~ContactModel() {}
    virtual const char* name() const = 0;
    /// the nb2_contact_model this model runs as on the device, or -1 for a model the device does not carry
    virtual int device_model() const { return -1; }
};
/// src/solver/signorini_coulomb_pyramid_model.rs: unilateral normal row + two friction-pyramid rows per contact
struct SignoriniCoulombPyramidModel : ContactModel {
    const char* name() const override { return "SignoriniCoulombPyramidModel"; }
    int device_model() const override { return NB2_CONTACT_SIGNORINI_COULOMB_PYRAMID; }
};
/// src/solver/signorini_model.rs:200-298: frictionless, active contacts only
struct SignoriniModel : ContactModel {
    const char* name() const override { return "SignoriniModel"; }
    int device_model() const override { return NB2_CONTACT_SIGNORINI; }
};

// counters/mod.rs (solver stages, milliseconds)
struct Counters {
    bool enabled = false;
    float assembly_time = 0.f, velocity_resolution_time = 0.f, velocity_update_time = 0.f, position_resolution_time = 0.f,
          solver_time = 0.f;
    size_t nconstraints = 0;
    void enable() { enabled = true; }
};

enum class SolverMode { ReferenceOrder = NB2_MODE_REFERENCE_ORDER, Coloured = NB2_MODE_COLOURED };

// ActivationManager (src/detection/activation_manager.rs:11-45): the island building and the sleeping
// rules themselves run on the device (nb2_update_activation); the host side keeps the mixing factor and
// the deferred activation requests.
struct ActivationManager {
    float mix_factor = 0.01f;  // mechanical_world.rs:80
    std::vector<int32_t> to_activate;
    void deferred_activate(DefaultBodyHandle h) { to_activate.push_back((int32_t)h); }
};

// ---------------------------------------------------------------------------------------------
// moreau_jean_solver.rs:14-90
class MoreauJeanSolver {
    nb2_context* ctx_ = nullptr;
    std::unique_ptr<ContactModel> contact_model_;
    std::vector<nb2_body> body_stage_;
    std::vector<nb2_body_state> state_stage_;
    std::vector<nb2_activation> act_stage_;
    std::vector<nb2_manifold> manifold_stage_;
    std::vector<nb2_contact> contact_stage_;
    std::vector<nb2_joint> joint_stage_;
    std::vector<nb2_multibody> mb_stage_;
    std::vector<nb2_mb_link> link_stage_;

    void check(int rc) const {
        if (rc != NB2_OK) throw SolverError(rc, nb2_last_error(ctx_) ? nb2_last_error(ctx_) : nb2_error_string(rc));
    }

  public:
    SolverMode mode = SolverMode::Coloured;
    Vector3 gravity{0.f, -9.81f, 0.f};

    explicit MoreauJeanSolver(std::unique_ptr<ContactModel> contact_model, int device = 0) : contact_model_(std::move(contact_model)) {
        int rc = nb2_create(device, nullptr, &ctx_);
        if (rc != NB2_OK) throw SolverError(rc, nb2_last_error(nullptr));
    }
    ~MoreauJeanSolver() { if (ctx_) nb2_destroy(ctx_); }
    MoreauJeanSolver(const MoreauJeanSolver&) = delete;
    MoreauJeanSolver& operator=(const MoreauJeanSolver&) = delete;

    /// moreau_jean_solver.rs:42-44.  A new model forgets the cached impulses.
    void set_contact_model(std::unique_ptr<ContactModel> model) {
        if (model->device_model() < 0)
            throw SolverError(NB2_ERR_UNSUPPORTED, "a user-defined ContactModel cannot run on the device: the two models "
                                                   "nphysics ships (SignoriniCoulombPyramidModel, SignoriniModel) can");
        check(nb2_set_contact_model(ctx_, model->device_model()));
        contact_model_ = std::move(model);
        check(nb2_clear_impulse_cache(ctx_));
    }

    /// MoreauJeanSolver::step (moreau_jean_solver.rs:47-61).  `island` / `island_joints` are implied: every
    /// dynamic body and every unbroken joint with a dynamic side (mechanical_world.rs:264-279).
    /// `activation`: the world's ActivationManager, or null for a bare solver step (every body awake).
    void step(Counters& counters, DefaultBodySet& bodies, DefaultJointConstraintSet& joints,
              const std::vector<ColliderContactManifold>& manifolds, const IntegrationParameters& parameters,
              ActivationManager* activation = nullptr) {
        upload_sets(counters, bodies, joints, parameters, activation);
        manifold_stage_.clear();
        contact_stage_.clear();
        for (const ColliderContactManifold& m : manifolds) {
            nb2_manifold rec = m.manifold;
            rec.first_contact = (uint32_t)contact_stage_.size();
            rec.num_contacts = (uint32_t)m.contacts.size();
            manifold_stage_.push_back(rec);
            contact_stage_.insert(contact_stage_.end(), m.contacts.begin(), m.contacts.end());
        }
        check(nb2_upload_manifolds(ctx_, manifold_stage_.data(), (uint32_t)manifold_stage_.size(), contact_stage_.data(),
                                   (uint32_t)contact_stage_.size()));
        run_and_download(counters, bodies, joints, activation);
        counters.nconstraints = 3 * contact_stage_.size();  // set_nconstraints (:74-76), contact rows
    }
    /// The same step with the contacts produced on the device from the collider set (cuboids):
    /// the collision-detection half of MechanicalWorld::step (mechanical_world.rs:241-262) included.
    void step(Counters& counters, GeometricalWorld& gworld, DefaultBodySet& bodies, DefaultColliderSet& colliders,
              DefaultJointConstraintSet& joints, const IntegrationParameters& parameters,
              ActivationManager* activation = nullptr) {
        const bool new_bodies = upload_sets(counters, bodies, joints, parameters, activation);
        bool broad_phase = gworld.broad_phase_interval <= 1 || gworld.steps_ % gworld.broad_phase_interval == 0;
        if (new_bodies || colliders.dirty_) {  // a new body set drops the colliders on the device
            collider_stage_.resize(colliders.colliders_.size());
            prediction_ = 0.f;
            for (size_t i = 0; i < collider_stage_.size(); ++i) {
                collider_stage_[i] = colliders.colliders_[i].rec_;
                prediction_ = std::fmax(prediction_, colliders.colliders_[i].linear_prediction_);
            }
            check(nb2_upload_colliders(ctx_, collider_stage_.data(), (uint32_t)collider_stage_.size()));
            colliders.dirty_ = false;
            broad_phase = true;
        }
        if (broad_phase) check(nb2_detect_pairs(ctx_, prediction_, gworld.search_radius, 0u, &gworld.n_pairs));
        ++gworld.steps_;
        check(nb2_generate_manifolds(ctx_));
        run_and_download(counters, bodies, joints, activation);
        counters.nconstraints = 12 * (size_t)gworld.n_pairs;  // an upper bound (<= 4 contacts x 3 rows per pair); stats() has the count
    }

  private:
    std::vector<nb2_collider> collider_stage_;
    float prediction_ = 0.001f;
    /// bodies, multibodies, joints, parameters: whatever changed on the host since the last step.  True when the
    /// body set went up again (the device then holds no joints, colliders or pairs of the old set).
    bool upload_sets(Counters& counters, DefaultBodySet& bodies, DefaultJointConstraintSet& joints,
                     const IntegrationParameters& parameters, ActivationManager* activation) {
        bool new_bodies = false;
        nb2_params p = parameters.to_abi(gravity);
        check(nb2_set_params(ctx_, &p));
        check(nb2_enable_timers(ctx_, counters.enabled ? 1 : 0));
        if (bodies.dirty_) {
            body_stage_.resize(bodies.bodies_.size());
            for (size_t i = 0; i < body_stage_.size(); ++i) body_stage_[i] = bodies.bodies_[i].rec_;
            check(nb2_upload_bodies(ctx_, body_stage_.data(), (uint32_t)body_stage_.size()));
            if (activation) {
                act_stage_.resize(bodies.bodies_.size());
                for (size_t i = 0; i < act_stage_.size(); ++i) act_stage_[i] = bodies.bodies_[i].act_;
                check(nb2_upload_activation(ctx_, act_stage_.data(), (uint32_t)act_stage_.size()));
            }
            mb_stage_.clear();
            link_stage_.clear();
            for (size_t m = 0; m < bodies.multibodies_.size(); ++m) {
                const Multibody& mb = bodies.multibodies_[m];
                nb2_multibody rec;
                rec.first_link = (uint32_t)link_stage_.size();
                rec.n_links = (uint32_t)mb.links_.size();
                rec.flags = mb.gravity_enabled_ ? NB2_BODY_FLAG_GRAVITY : 0u;
                rec.reserved = 0;
                mb_stage_.push_back(rec);
                for (size_t i = 0; i < mb.links_.size(); ++i) {
                    nb2_mb_link l = mb.links_[i];
                    l.multibody = (int32_t)m;
                    l.body = (int32_t)(bodies.mb_first_body_[m] + i);
                    link_stage_.push_back(l);
                }
            }
            check(nb2_upload_multibodies(ctx_, mb_stage_.data(), (uint32_t)mb_stage_.size(), link_stage_.data(), (uint32_t)link_stage_.size()));
            bodies.dirty_ = false;
            joints.dirty_ = true;  // a new body set drops the joints on device
            new_bodies = true;
        }
        if (joints.dirty_) {
            joint_stage_.resize(joints.joints_.size());
            for (size_t i = 0; i < joint_stage_.size(); ++i) joint_stage_[i] = joints.joints_[i].record();
            check(nb2_upload_joints(ctx_, joint_stage_.data(), (uint32_t)joint_stage_.size()));
            joints.dirty_ = false;
        }
        return new_bodies;
    }
    /// activation update, the step, and the outputs written in place into the bodies / joints, like the reference
    void run_and_download(Counters& counters, DefaultBodySet& bodies, DefaultJointConstraintSet& joints,
                          ActivationManager* activation) {
        if (activation) {  // ActivationManager::update sits between the narrow phase and the solver (mechanical_world.rs:265-272)
            check(nb2_update_activation(ctx_, activation->mix_factor, activation->to_activate.data(),
                                        (uint32_t)activation->to_activate.size()));
            activation->to_activate.clear();
        }
        check(nb2_step(ctx_, (int)mode));
        // outputs are written in place into the bodies / joints, like the reference
        state_stage_.resize(bodies.bodies_.size());
        if (activation) {
            check(nb2_download_activation(ctx_, act_stage_.data(), (uint32_t)act_stage_.size()));
            for (size_t i = 0; i < act_stage_.size(); ++i) bodies.bodies_[i].act_ = act_stage_[i];
        }
        check(nb2_download_body_states(ctx_, state_stage_.data(), 0, (uint32_t)state_stage_.size()));
        for (size_t i = 0; i < state_stage_.size(); ++i) {
            std::memcpy(bodies.bodies_[i].rec_.position, state_stage_[i].position, sizeof(float) * 7);
            std::memcpy(bodies.bodies_[i].rec_.velocity, state_stage_[i].velocity, sizeof(float) * 6);
        }
        if (!link_stage_.empty()) {  // joint coordinates, generalized velocities, cached impulses; link poses came with the body states
            check(nb2_download_multibody_links(ctx_, link_stage_.data(), (uint32_t)link_stage_.size()));
            size_t k = 0;
            for (size_t m = 0; m < bodies.multibodies_.size(); ++m) {
                Multibody& mb = bodies.multibodies_[m];
                for (size_t i = 0; i < mb.links_.size(); ++i, ++k) {
                    mb.links_[i] = link_stage_[k];
                    mb.parts_[i] = bodies.bodies_[bodies.mb_first_body_[m] + i].rec_;
                }
            }
        }
        if (!joint_stage_.empty()) {
            check(nb2_download_joints(ctx_, joint_stage_.data(), (uint32_t)joint_stage_.size()));
            for (size_t i = 0; i < joint_stage_.size(); ++i) joints.joints_[i].record_mut() = joint_stage_[i];
        }
        check(nb2_synchronize(ctx_));
        if (counters.enabled) {
            float t[8];
            check(nb2_get_timers(ctx_, t));
            counters.assembly_time = t[0];
            counters.velocity_resolution_time = t[1];
            counters.velocity_update_time = t[2];
            counters.position_resolution_time = t[3];
            counters.solver_time = t[4];
        }
    }

  public:
    nb2_stats stats() {
        nb2_stats s;
        check(nb2_get_stats(ctx_, &s));
        return s;
    }
};

// mechanical_world.rs:55-68, 182-396 -- the solver part of the step; collision detection (the
// manifold producer) stays with the caller.
class MechanicalWorld {
  public:
    Counters counters;
    MoreauJeanSolver solver;
    IntegrationParameters integration_parameters;
    Vector3 gravity;

    explicit MechanicalWorld(const Vector3& g, int device = 0)
        : solver(std::unique_ptr<ContactModel>(new SignoriniCoulombPyramidModel()), device), gravity(g) {}
    void set_timestep(float dt) { integration_parameters.set_dt(dt); }
    float timestep() const { return integration_parameters.dt(); }
    ActivationManager activation_manager;  // mechanical_world.rs:66,80
    void step(DefaultBodySet& bodies, DefaultJointConstraintSet& joints, const std::vector<ColliderContactManifold>& manifolds) {
        solver.gravity = gravity;
        solver.step(counters, bodies, joints, manifolds, integration_parameters, &activation_manager);
    }
    /// mechanical_world.rs:182-188: `step(&mut geometrical_world, &mut bodies, &mut colliders, &mut joint_constraints, ..)`
    /// with the collision detection on the device (cuboid colliders).
    void step(GeometricalWorld& geometrical_world, DefaultBodySet& bodies, DefaultColliderSet& colliders,
              DefaultJointConstraintSet& joints) {
        solver.gravity = gravity;
        solver.step(counters, geometrical_world, bodies, colliders, joints, integration_parameters, &activation_manager);
    }
};

}  // namespace nphysics
