//! Raw bindings to `include/nphysics_b200.h` (ABI version 1).  Field for field, in header order.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_void};

pub const NB2_OK: i32 = 0;
pub const NB2_ERR_INVALID_ARGUMENT: i32 = -1;
pub const NB2_ERR_NO_DEVICE: i32 = -2;
pub const NB2_ERR_CUDA: i32 = -3;
pub const NB2_ERR_OUT_OF_MEMORY: i32 = -4;
pub const NB2_ERR_BAD_INDEX: i32 = -5;
pub const NB2_ERR_UNSUPPORTED: i32 = -6;
pub const NB2_ERR_TOO_MANY_COLOURS: i32 = -7;
pub const NB2_ERR_NOT_READY: i32 = -8;
pub const NB2_ERR_NON_FINITE: i32 = -9;

pub const NB2_BODY_DISABLED: u32 = 0;
pub const NB2_BODY_STATIC: u32 = 1;
pub const NB2_BODY_DYNAMIC: u32 = 2;
pub const NB2_BODY_KINEMATIC: u32 = 3;
pub const NB2_BODY_FLAG_GRAVITY: u32 = 1;

pub const NB2_GEOM_POINT: u8 = 0;
pub const NB2_GEOM_LINE: u8 = 1;
pub const NB2_GEOM_PLANE: u8 = 2;

pub const NB2_JOINT_BALL: u32 = 0;
pub const NB2_JOINT_REVOLUTE: u32 = 1;
pub const NB2_JOINT_PRISMATIC: u32 = 2;
pub const NB2_JOINT_UNIVERSAL: u32 = 3;
pub const NB2_JOINT_PLANAR: u32 = 4;
pub const NB2_JOINT_RECTANGULAR: u32 = 5;
pub const NB2_JOINT_PIN_SLOT: u32 = 6;
pub const NB2_JOINT_CYLINDRICAL: u32 = 7;
pub const NB2_JOINT_FIXED: u32 = 8;
pub const NB2_JOINT_CARTESIAN: u32 = 9;
pub const NB2_JOINT_FLAG_MIN_OFFSET: u32 = 1;
pub const NB2_JOINT_FLAG_MAX_OFFSET: u32 = 2;

pub const NB2_CONTACT_SIGNORINI_COULOMB_PYRAMID: i32 = 0;
pub const NB2_CONTACT_SIGNORINI: i32 = 1;
pub const NB2_MODE_REFERENCE_ORDER: i32 = 0;
pub const NB2_MODE_COLOURED: i32 = 1;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_params {
    pub dt: f32,
    pub erp: f32,
    pub warmstart_coeff: f32,
    pub restitution_velocity_threshold: f32,
    pub allowed_linear_error: f32,
    pub allowed_angular_error: f32,
    pub max_linear_correction: f32,
    pub max_angular_correction: f32,
    pub max_stabilization_multiplier: f32,
    pub max_velocity_iterations: u32,
    pub max_position_iterations: u32,
    pub max_ccd_position_iterations: u32,
    pub max_ccd_substeps: u32,
    pub gravity: [f32; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_contact_update {
    pub world1: [f32; 3],
    pub world2: [f32; 3],
    pub normal: [f32; 3],
    pub depth: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_collider {
    pub half_extents: [f32; 3],
    pub margin: f32,
    pub translation_wrt_body: [f32; 3],
    pub friction: f32,
    pub rotation_wrt_body: [f32; 4],
    pub restitution: f32,
    pub body: i32,
    pub friction_mode: u8,
    pub restitution_mode: u8,
    pub pad_: [u8; 2],
    pub flags: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_body {
    pub position: [f32; 7],
    pub velocity: [f32; 6],
    pub local_com: [f32; 3],
    pub mass: f32,
    pub local_inertia: [f32; 9],
    pub external_forces: [f32; 6],
    pub linear_damping: f32,
    pub angular_damping: f32,
    pub max_linear_velocity: f32,
    pub max_angular_velocity: f32,
    pub jacobian_mask: [f32; 6],
    pub status: u32,
    pub flags: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_body_state {
    pub position: [f32; 7],
    pub velocity: [f32; 6],
}

/// ActivationStatus (src/object/body.rs:65-125): `threshold < 0` stands for `None`, `energy == 0` for asleep.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_activation {
    pub threshold: f32,
    pub energy: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_manifold {
    pub body1: i32,
    pub body2: i32,
    pub first_contact: u32,
    pub num_contacts: u32,
    pub margin1: f32,
    pub margin2: f32,
    pub friction: f32,
    pub restitution: f32,
    pub surface_velocity: [f32; 3],
    pub coll1_wrt_body: [f32; 7],
    pub coll2_wrt_body: [f32; 7],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_contact {
    pub world1: [f32; 3],
    pub world2: [f32; 3],
    pub normal: [f32; 3],
    pub depth: f32,
    pub key: u64,
    pub local1: [f32; 3],
    pub local2: [f32; 3],
    pub dir1: [f32; 3],
    pub dir2: [f32; 3],
    pub dilation1: f32,
    pub dilation2: f32,
    pub geom1: u8,
    pub geom2: u8,
    pub pad_: [u8; 6],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_joint {
    pub type_: u32,
    pub body1: i32,
    pub body2: i32,
    pub flags: u32,
    pub anchor1: [f32; 3],
    pub anchor2: [f32; 3],
    pub axis1: [f32; 3],
    pub axis2: [f32; 3],
    pub axis3: [f32; 3],
    pub ref_frame1: [f32; 4],
    pub ref_frame2: [f32; 4],
    pub angle: f32,
    pub min_offset: f32,
    pub max_offset: f32,
    pub break_force_squared: f32,
    pub break_torque_squared: f32,
    pub impulses: [f32; 7],
    pub broken: u32,
}

pub const NB2_BODY_MULTIBODY_LINK: u32 = 4;
pub const NB2_MBJ_FREE: u32 = 0;
pub const NB2_MBJ_BALL: u32 = 1;
pub const NB2_MBJ_REVOLUTE: u32 = 2;
pub const NB2_MBJ_PRISMATIC: u32 = 3;
pub const NB2_MBJ_FIXED: u32 = 4;
pub const NB2_MBJ_FLAG_MIN: u32 = 1;
pub const NB2_MBJ_FLAG_MAX: u32 = 2;
pub const NB2_MBJ_FLAG_MOTOR: u32 = 4;

/// One link of a reduced-coordinate multibody (`MultibodyLink` + its `Joint`, src/object/multibody_link.rs).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct nb2_mb_link {
    pub multibody: i32,
    pub parent: i32,
    pub joint_type: u32,
    pub flags: u32,
    pub body: i32,
    pub parent_shift: [f32; 3],
    pub body_shift: [f32; 3],
    pub axis: [f32; 3],
    pub coords: [f32; 7],
    pub velocity: [f32; 6],
    pub damping: [f32; 6],
    pub min_pos: f32,
    pub max_pos: f32,
    pub motor_velocity: f32,
    pub motor_max_velocity: f32,
    pub motor_max_force: f32,
    pub impulses: [f32; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct nb2_multibody {
    pub first_link: u32,
    pub n_links: u32,
    pub flags: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct nb2_stats {
    pub n_bodies: u32,
    pub n_dynamic_bodies: u32,
    pub n_manifolds: u32,
    pub n_contacts: u32,
    pub n_joints: u32,
    pub n_rows_two_body: u32,
    pub n_rows_ground: u32,
    pub n_phases_velocity: u32,
    pub n_phases_position: u32,
    pub n_broken_joints: u32,
    pub non_finite: u32,
    pub schedule_verdict: u32,
    pub residual_max: f32,
    pub residual_rms: f32,
    pub max_penetration: f32,
    pub kinetic_energy: f32,
    pub t_assembly_ms: f32,
    pub t_velocity_resolution_ms: f32,
    pub t_velocity_update_ms: f32,
    pub t_position_resolution_ms: f32,
    pub t_step_ms: f32,
    pub pad2_: f32,
}

pub enum nb2_context {}

extern "C" {
    pub fn nb2_abi_version() -> i32;
    pub fn nb2_error_string(err: i32) -> *const c_char;
    pub fn nb2_default_params(out: *mut nb2_params) -> i32;
    pub fn nb2_sizeof(which: i32) -> i32;
    pub fn nb2_combine_materials(
        friction1: f32, friction_mode1: i32, restitution1: f32, restitution_mode1: i32, surface_velocity1: *const f32,
        friction2: f32, friction_mode2: i32, restitution2: f32, restitution_mode2: i32, surface_velocity2: *const f32,
        out_friction: *mut f32, out_restitution: *mut f32, out_surface_velocity3: *mut f32,
    ) -> i32;
    pub fn nb2_create(device: i32, stream: *mut c_void, out_ctx: *mut *mut nb2_context) -> i32;
    pub fn nb2_destroy(ctx: *mut nb2_context) -> i32;
    pub fn nb2_last_error(ctx: *const nb2_context) -> *const c_char;
    pub fn nb2_set_params(ctx: *mut nb2_context, params: *const nb2_params) -> i32;
    pub fn nb2_get_params(ctx: *const nb2_context, out: *mut nb2_params) -> i32;
    pub fn nb2_enable_timers(ctx: *mut nb2_context, enabled: i32) -> i32;
    pub fn nb2_set_schedule_cache(ctx: *mut nb2_context, enabled: i32) -> i32;
    pub fn nb2_set_contact_layout(ctx: *mut nb2_context, layout: i32) -> i32;
    pub fn nb2_upload_bodies(ctx: *mut nb2_context, bodies: *const nb2_body, n: u32) -> i32;
    pub fn nb2_upload_body_states(ctx: *mut nb2_context, states: *const nb2_body_state, first: u32, n: u32) -> i32;
    pub fn nb2_upload_manifolds(ctx: *mut nb2_context, manifolds: *const nb2_manifold, n_manifolds: u32,
                                contacts: *const nb2_contact, n_contacts: u32) -> i32;
    pub fn nb2_upload_joints(ctx: *mut nb2_context, joints: *const nb2_joint, n_joints: u32) -> i32;
    pub fn nb2_upload_multibodies(ctx: *mut nb2_context, multibodies: *const nb2_multibody, n_multibodies: u32,
                                  links: *const nb2_mb_link, n_links: u32) -> i32;
    pub fn nb2_download_multibody_links(ctx: *mut nb2_context, out: *mut nb2_mb_link, n_links: u32) -> i32;
    pub fn nb2_clear_impulse_cache(ctx: *mut nb2_context) -> i32;
    pub fn nb2_upload_activation(ctx: *mut nb2_context, activation: *const nb2_activation, n: u32) -> i32;
    pub fn nb2_update_activation(ctx: *mut nb2_context, mix_factor: f32, to_activate: *const i32, n_to_activate: u32) -> i32;
    pub fn nb2_download_activation(ctx: *mut nb2_context, out: *mut nb2_activation, n: u32) -> i32;
    pub fn nb2_step(ctx: *mut nb2_context, mode: i32) -> i32;
    pub fn nb2_step_ccd(ctx: *mut nb2_context, mode: i32) -> i32;
    pub fn nb2_synchronize(ctx: *mut nb2_context) -> i32;
    pub fn nb2_download_body_states(ctx: *mut nb2_context, out: *mut nb2_body_state, first: u32, n: u32) -> i32;
    pub fn nb2_download_contact_impulses(ctx: *mut nb2_context, out3: *mut f32, n_contacts: u32) -> i32;
    pub fn nb2_download_joints(ctx: *mut nb2_context, out: *mut nb2_joint, n_joints: u32) -> i32;
    pub fn nb2_get_stats(ctx: *mut nb2_context, out: *mut nb2_stats) -> i32;
    pub fn nb2_get_timers(ctx: *mut nb2_context, out8: *mut f32) -> i32;
    pub fn nb2_launch_count(ctx: *const nb2_context, out: *mut u64) -> i32;
    pub fn nb2_update_contacts(ctx: *mut nb2_context, updates: *const nb2_contact_update, n_contacts: u32) -> i32;
    pub fn nb2_upload_colliders(ctx: *mut nb2_context, colliders: *const nb2_collider, n_colliders: u32) -> i32;
    pub fn nb2_detect_pairs(ctx: *mut nb2_context, linear_prediction: f32, search_radius: f32, flip_permille: u32,
                            out_pairs: *mut u32) -> i32;
    pub fn nb2_generate_manifolds(ctx: *mut nb2_context) -> i32;
    pub fn nb2_download_manifolds(ctx: *mut nb2_context, out_manifolds: *mut nb2_manifold, manifold_capacity: u32,
                                  out_contacts: *mut nb2_contact, contact_capacity: u32, out_n_manifolds: *mut u32,
                                  out_n_contacts: *mut u32) -> i32;
    pub fn nb2_set_contact_model(ctx: *mut nb2_context, model: i32) -> i32;
    pub fn nb2_label_islands(ctx: *mut nb2_context, out_labels: *mut i32, out_rows: *mut u32, n: u32) -> i32;
    pub fn nb2_download_schedule(ctx: *mut nb2_context, out_phase: *mut i32, out_body1: *mut i32, out_body2: *mut i32,
                                 capacity: u32, out_n: *mut u32) -> i32;
}

#[cfg(test)]
mod tests {
    use super::*;
    #[test]
    fn layouts_match_the_header() {
        unsafe {
            assert_eq!(nb2_sizeof(0) as usize, core::mem::size_of::<nb2_params>());
            assert_eq!(nb2_sizeof(1) as usize, core::mem::size_of::<nb2_body>());
            assert_eq!(nb2_sizeof(2) as usize, core::mem::size_of::<nb2_body_state>());
            assert_eq!(nb2_sizeof(3) as usize, core::mem::size_of::<nb2_manifold>());
            assert_eq!(nb2_sizeof(4) as usize, core::mem::size_of::<nb2_contact>());
            assert_eq!(nb2_sizeof(5) as usize, core::mem::size_of::<nb2_joint>());
            assert_eq!(nb2_sizeof(6) as usize, core::mem::size_of::<nb2_stats>());
        }
    }
}
