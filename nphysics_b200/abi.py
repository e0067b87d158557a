"""numpy / ctypes mirror of include/nphysics_b200.h (the C-ABI structs).

Records travel as numpy structured arrays whose layout is checked against
``nb2_sizeof`` when the library is loaded.  Field names are the header's.
"""
import ctypes

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)

# nb2_error
OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_NO_DEVICE = -2
ERR_CUDA = -3
ERR_OUT_OF_MEMORY = -4
ERR_BAD_INDEX = -5
ERR_UNSUPPORTED = -6
ERR_TOO_MANY_COLOURS = -7
ERR_NOT_READY = -8
ERR_NON_FINITE = -9

# nb2_body_status
BODY_DISABLED, BODY_STATIC, BODY_DYNAMIC, BODY_KINEMATIC, BODY_MULTIBODY_LINK = 0, 1, 2, 3, 4
BODY_FLAG_GRAVITY = 1

# nb2_kinematic_geom
GEOM_POINT, GEOM_LINE, GEOM_PLANE = 0, 1, 2

# nb2_joint_type
(JOINT_BALL, JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_UNIVERSAL, JOINT_PLANAR, JOINT_RECTANGULAR,
 JOINT_PIN_SLOT, JOINT_CYLINDRICAL, JOINT_FIXED, JOINT_CARTESIAN) = range(10)
JOINT_FLAG_MIN_OFFSET, JOINT_FLAG_MAX_OFFSET = 1, 2

# nb2_mb_joint_type (reduced-coordinate joints of a multibody link)
MBJ_FREE, MBJ_BALL, MBJ_REVOLUTE, MBJ_PRISMATIC, MBJ_FIXED = range(5)
MBJ_FLAG_MIN, MBJ_FLAG_MAX, MBJ_FLAG_MOTOR = 1, 2, 4
MBJ_NDOFS = {MBJ_FREE: 6, MBJ_BALL: 3, MBJ_REVOLUTE: 1, MBJ_PRISMATIC: 1, MBJ_FIXED: 0}
MB_MAX_DOFS = 64

# nb2_step_mode
MODE_REFERENCE_ORDER, MODE_COLOURED = 0, 1

# nb2_contact_model
CONTACT_SIGNORINI_COULOMB_PYRAMID, CONTACT_SIGNORINI = 0, 1

f4, u4, i4, u8, u1 = np.float32, np.uint32, np.int32, np.uint64, np.uint8

params_dtype = np.dtype([
    ("dt", f4), ("erp", f4), ("warmstart_coeff", f4), ("restitution_velocity_threshold", f4),
    ("allowed_linear_error", f4), ("allowed_angular_error", f4), ("max_linear_correction", f4),
    ("max_angular_correction", f4), ("max_stabilization_multiplier", f4),
    ("max_velocity_iterations", u4), ("max_position_iterations", u4),
    ("max_ccd_position_iterations", u4), ("max_ccd_substeps", u4), ("gravity", f4, 3)], align=True)

body_dtype = np.dtype([
    ("position", f4, 7), ("velocity", f4, 6), ("local_com", f4, 3), ("mass", f4),
    ("local_inertia", f4, 9), ("external_forces", f4, 6), ("linear_damping", f4),
    ("angular_damping", f4), ("max_linear_velocity", f4), ("max_angular_velocity", f4),
    ("jacobian_mask", f4, 6), ("status", u4), ("flags", u4)], align=True)

body_state_dtype = np.dtype([("position", f4, 7), ("velocity", f4, 6)], align=True)

manifold_dtype = np.dtype([
    ("body1", i4), ("body2", i4), ("first_contact", u4), ("num_contacts", u4), ("margin1", f4),
    ("margin2", f4), ("friction", f4), ("restitution", f4), ("surface_velocity", f4, 3),
    ("coll1_wrt_body", f4, 7), ("coll2_wrt_body", f4, 7)], align=True)

contact_dtype = np.dtype([
    ("world1", f4, 3), ("world2", f4, 3), ("normal", f4, 3), ("depth", f4), ("key", u8),
    ("local1", f4, 3), ("local2", f4, 3), ("dir1", f4, 3), ("dir2", f4, 3), ("dilation1", f4),
    ("dilation2", f4), ("geom1", u1), ("geom2", u1), ("pad_", u1, 6)], align=True)

joint_dtype = np.dtype([
    ("type", u4), ("body1", i4), ("body2", i4), ("flags", u4), ("anchor1", f4, 3), ("anchor2", f4, 3),
    ("axis1", f4, 3), ("axis2", f4, 3), ("axis3", f4, 3), ("ref_frame1", f4, 4), ("ref_frame2", f4, 4),
    ("angle", f4), ("min_offset", f4), ("max_offset", f4), ("break_force_squared", f4),
    ("break_torque_squared", f4), ("impulses", f4, 7), ("broken", u4)], align=True)

stats_dtype = np.dtype([
    ("n_bodies", u4), ("n_dynamic_bodies", u4), ("n_manifolds", u4), ("n_contacts", u4), ("n_joints", u4),
    ("n_rows_two_body", u4), ("n_rows_ground", u4), ("n_phases_velocity", u4), ("n_phases_position", u4),
    ("n_broken_joints", u4), ("non_finite", u4), ("schedule_verdict", u4), ("residual_max", f4), ("residual_rms", f4),
    ("max_penetration", f4), ("kinetic_energy", f4), ("t_assembly_ms", f4),
    ("t_velocity_resolution_ms", f4), ("t_velocity_update_ms", f4), ("t_position_resolution_ms", f4),
    ("t_step_ms", f4), ("pad2_", f4)], align=True)

# index used by nb2_sizeof(which)
# nb2_activation: ActivationStatus {threshold (< 0 = None), energy (0 = asleep)} (body.rs:65-125)
activation_dtype = np.dtype([("threshold", f4), ("energy", f4)], align=True)
DEFAULT_SLEEP_THRESHOLD = 0.01  # ActivationStatus::default_threshold (body.rs:72-74)


def new_activation(n, threshold=DEFAULT_SLEEP_THRESHOLD):
    """ActivationStatus::new_active for n bodies (threshold=None -> bodies that never sleep)."""
    a = np.zeros(n, dtype=activation_dtype)
    a["threshold"] = -1.0 if threshold is None else threshold
    a["energy"] = 0.04 if threshold is None else 4.0 * threshold
    return a


# nb2_contact_update: the per-step part of a TrackedContact (the leading 40 bytes of nb2_contact)
contact_update_dtype = np.dtype([("world1", f4, 3), ("world2", f4, 3), ("normal", f4, 3), ("depth", f4)], align=True)

# nb2_collider: cuboid collider of the device manifold producer
collider_dtype = np.dtype([
    ("half_extents", f4, 3), ("margin", f4), ("translation_wrt_body", f4, 3), ("friction", f4),
    ("rotation_wrt_body", f4, 4), ("restitution", f4), ("body", i4), ("friction_mode", u1),
    ("restitution_mode", u1), ("pad_", u1, 2), ("flags", u4)], align=True)

# nb2_mb_link / nb2_multibody: reduced-coordinate multibodies
mb_link_dtype = np.dtype([
    ("multibody", i4), ("parent", i4), ("joint_type", u4), ("flags", u4), ("body", i4),
    ("parent_shift", f4, 3), ("body_shift", f4, 3), ("axis", f4, 3), ("coords", f4, 7), ("velocity", f4, 6),
    ("damping", f4, 6), ("min_pos", f4), ("max_pos", f4), ("motor_velocity", f4), ("motor_max_velocity", f4),
    ("motor_max_force", f4), ("impulses", f4, 3)], align=True)
multibody_dtype = np.dtype([("first_link", u4), ("n_links", u4), ("flags", u4), ("reserved", u4)], align=True)

SIZEOF_ORDER = [params_dtype, body_dtype, body_state_dtype, manifold_dtype, contact_dtype, joint_dtype,
                stats_dtype, activation_dtype, contact_update_dtype, collider_dtype, mb_link_dtype, multibody_dtype]
EXPECTED_SIZES = [64, 176, 52, 100, 112, 160, 88, 8, 40, 64, 164, 16]

for _d, _s in zip(SIZEOF_ORDER, EXPECTED_SIZES):
    assert _d.itemsize == _s, (_d, _d.itemsize, _s)


def ptr(arr):
    """void* to a C-contiguous numpy array (or None)."""
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(ctypes.c_void_p)


def default_params():
    """IntegrationParameters::default() (integration_parameters.rs:169-189) + gravity of the
    examples (examples3d/pyramid3.rs:21)."""
    p = np.zeros((), dtype=params_dtype)
    p["dt"] = 1.0 / 60.0
    p["erp"] = 0.2
    p["warmstart_coeff"] = 1.0
    p["restitution_velocity_threshold"] = 1.0
    p["allowed_linear_error"] = 0.001
    p["allowed_angular_error"] = 0.001
    p["max_linear_correction"] = 0.2
    p["max_angular_correction"] = 0.2
    p["max_stabilization_multiplier"] = 0.2
    p["max_velocity_iterations"] = 8
    p["max_position_iterations"] = 3
    p["max_ccd_position_iterations"] = 10
    p["max_ccd_substeps"] = 1
    p["gravity"] = (0.0, -9.81, 0.0)
    return p


def new_bodies(n):
    """n zeroed body records with the RigidBody::new defaults (rigid_body.rs:55-85)."""
    b = np.zeros(n, dtype=body_dtype)
    b["position"][:, 6] = 1.0
    b["max_linear_velocity"] = FLT_MAX
    b["max_angular_velocity"] = FLT_MAX
    b["jacobian_mask"] = 1.0
    b["status"] = BODY_DYNAMIC
    b["flags"] = BODY_FLAG_GRAVITY
    return b


def new_joints(n, jtype=JOINT_BALL):
    j = np.zeros(n, dtype=joint_dtype)
    j["type"] = jtype
    j["ref_frame1"][:, 3] = 1.0
    j["ref_frame2"][:, 3] = 1.0
    j["axis1"][:, 0] = 1.0
    j["axis2"][:, 0] = 1.0
    j["axis3"][:, 0] = 1.0
    j["break_force_squared"] = FLT_MAX
    j["break_torque_squared"] = FLT_MAX
    return j


def new_colliders(n):
    c = np.zeros(n, dtype=collider_dtype)
    c["rotation_wrt_body"][:, 3] = 1.0
    c["margin"] = 0.01       # ColliderDesc::default_margin (collider.rs:457-479)
    c["friction"] = 0.5      # BasicMaterial::default (basic_material.rs:56-60)
    return c


def contact_updates_of(contacts):
    """The per-step part (nb2_contact_update) of full contact records."""
    u = np.zeros(len(contacts), dtype=contact_update_dtype)
    for f in ("world1", "world2", "normal", "depth"):
        u[f] = contacts[f]
    return u


def new_mb_links(n, joint_type=MBJ_BALL):
    """n multibody links with the joints' defaults: identity coordinates, Joint::default_damping (0.1 on the dofs
    of ball and revolute joints: ball_joint.rs:89-91, revolute_joint.rs:214-216), no limits, JointMotor::new()."""
    l = np.zeros(n, dtype=mb_link_dtype)
    l["parent"] = -1
    l["joint_type"] = joint_type
    l["body"] = -1
    l["axis"][:, 0] = 1.0
    l["motor_max_velocity"] = FLT_MAX
    l["motor_max_force"] = FLT_MAX
    set_mb_joint(l, slice(None), joint_type)
    return l


def set_mb_joint(links, idx, joint_type):
    """Sets the joint type of links[idx] together with its identity coordinates and default damping."""
    links["joint_type"][idx] = joint_type
    links["coords"][idx] = 0.0
    links["damping"][idx] = 0.0
    if joint_type in (MBJ_FREE, MBJ_FIXED):
        c = links["coords"][idx]
        c[..., 6] = 1.0
        links["coords"][idx] = c
    elif joint_type == MBJ_BALL:
        c = links["coords"][idx]
        c[..., 3] = 1.0
        links["coords"][idx] = c
        d = links["damping"][idx]
        d[..., :3] = 0.1
        links["damping"][idx] = d
    elif joint_type == MBJ_REVOLUTE:
        d = links["damping"][idx]
        d[..., 0] = 0.1
        links["damping"][idx] = d
