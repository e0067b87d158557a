//! Constraint-based joints -> `nb2_joint` records and back.
//!
//! The fields of nphysics' joint constraints are private and have no getters (e.g.
//! src/joint/ball_constraint.rs:14-24, src/joint/revolute_constraint.rs:36-50), so the impl blocks
//! below cannot live in an outside crate: INTEGRATION.md shows them added to nphysics itself (one impl
//! per `src/joint/*_constraint.rs`, where the fields are in scope) together with the `B200Joint` trait.
//! The impulse slots follow each type's `cache_impulses` verbatim (include/nphysics_b200.h, nb2_joint):
//! lin[0..3] | ang[3..6] | limit[6].
use nphysics3d::object::BodyHandle;
use nphysics_b200_sys as sys;

/// Implemented by every `*Constraint` type of nphysics (the patch of INTEGRATION.md section 3).
pub trait B200Joint<Handle: BodyHandle> {
    /// The record the solver uploads; `index_of` maps a body handle to its index in the uploaded body set.
    fn to_record(&self, index_of: &dyn Fn(Handle) -> i32) -> sys::nb2_joint;
    /// What `cache_impulses` would have stored (src/joint/ball_constraint.rs:132-144 and friends).
    fn store_solver_outputs(&mut self, rec: &sys::nb2_joint);
}

pub fn blank(type_: u32, body1: i32, body2: i32) -> sys::nb2_joint {
    let mut j: sys::nb2_joint = unsafe { std::mem::zeroed() };
    j.type_ = type_;
    j.body1 = body1;
    j.body2 = body2;
    j.axis1 = [1.0, 0.0, 0.0];
    j.axis2 = [1.0, 0.0, 0.0];
    j.axis3 = [1.0, 0.0, 0.0];
    j.ref_frame1 = [0.0, 0.0, 0.0, 1.0];
    j.ref_frame2 = [0.0, 0.0, 0.0, 1.0];
    j.break_force_squared = f32::MAX;
    j.break_torque_squared = f32::MAX;
    j
}

/// The in-crate impls, written against the private field names of nphysics3d 0.23.  Each block is pasted
/// next to the struct it reads (`use crate::b200::{B200Joint, blank, sys};`).
#[cfg(feature = "in-crate-patch")]
mod patch {
    use super::*;
    use nphysics3d::joint::*;

    fn v3(v: &nalgebra::Vector3<f32>) -> [f32; 3] { [v.x, v.y, v.z] }
    fn p3(p: &nalgebra::Point3<f32>) -> [f32; 3] { [p.x, p.y, p.z] }
    fn q4(q: &nalgebra::UnitQuaternion<f32>) -> [f32; 4] { [q.i, q.j, q.k, q.w] }

    macro_rules! common {
        ($s:ident, $ty:expr, $index_of:ident) => {{
            let mut j = blank($ty, $index_of($s.b1.0), $index_of($s.b2.0));
            j.anchor1 = p3(&$s.anchor1);
            j.anchor2 = p3(&$s.anchor2);
            j.broken = $s.broken as u32;
            j
        }};
    }

    impl<H: BodyHandle> B200Joint<H> for BallConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_BALL, index_of);
            j.break_force_squared = self.break_force_squared;
            j.impulses[..3].copy_from_slice(&v3(&self.impulses));
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.impulses = nalgebra::Vector3::new(r.impulses[0], r.impulses[1], r.impulses[2]);
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for RevoluteConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_REVOLUTE, index_of);
            j.axis1 = v3(&self.axis1);
            j.axis2 = v3(&self.axis2);
            j.break_force_squared = self.break_force_squared;
            j.break_torque_squared = self.break_torque_squared;
            j.impulses[..3].copy_from_slice(&v3(&self.lin_impulses));
            j.impulses[3..6].copy_from_slice(&v3(&self.ang_impulses));
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.lin_impulses = nalgebra::Vector3::new(r.impulses[0], r.impulses[1], r.impulses[2]);
            self.ang_impulses = nalgebra::Vector3::new(r.impulses[3], r.impulses[4], r.impulses[5]);
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for PrismaticConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_PRISMATIC, index_of);
            j.axis1 = v3(&self.axis1);
            j.break_force_squared = self.break_force_squared;
            j.break_torque_squared = self.break_torque_squared;
            if let Some(m) = self.min_offset { j.flags |= sys::NB2_JOINT_FLAG_MIN_OFFSET; j.min_offset = m; }
            if let Some(m) = self.max_offset { j.flags |= sys::NB2_JOINT_FLAG_MAX_OFFSET; j.max_offset = m; }
            j.impulses[..3].copy_from_slice(&v3(&self.lin_impulses));
            j.impulses[3..6].copy_from_slice(&v3(&self.ang_impulses));
            j.impulses[6] = self.limit_impulse;
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.lin_impulses = nalgebra::Vector3::new(r.impulses[0], r.impulses[1], r.impulses[2]);
            self.ang_impulses = nalgebra::Vector3::new(r.impulses[3], r.impulses[4], r.impulses[5]);
            self.limit_impulse = r.impulses[6];
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for UniversalConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_UNIVERSAL, index_of);
            j.axis1 = v3(&self.axis1);
            j.axis2 = v3(&self.axis2);
            j.angle = self.angle;
            j.break_force_squared = self.break_force_squared;
            j.break_torque_squared = self.break_torque_squared;
            j.impulses[..3].copy_from_slice(&v3(&self.lin_impulses));
            j.impulses[3] = self.ang_impulse;
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.lin_impulses = nalgebra::Vector3::new(r.impulses[0], r.impulses[1], r.impulses[2]);
            self.ang_impulse = r.impulses[3];
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for PlanarConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_PLANAR, index_of);
            j.axis1 = v3(&self.axis1);
            j.axis2 = v3(&self.axis2);
            j.break_force_squared = self.break_force_squared;
            j.break_torque_squared = self.break_torque_squared;
            j.impulses[0] = self.lin_impulse;
            j.impulses[3] = self.ang_impulses[0];
            j.impulses[4] = self.ang_impulses[1];
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.lin_impulse = r.impulses[0];
            self.ang_impulses = [r.impulses[3], r.impulses[4]];
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for RectangularConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_RECTANGULAR, index_of);
            j.axis1 = v3(&self.axis1);
            j.break_force_squared = self.break_force_squared;
            j.break_torque_squared = self.break_torque_squared;
            j.impulses[0] = self.lin_impulse;
            j.impulses[3..6].copy_from_slice(&v3(&self.ang_impulses));
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.lin_impulse = r.impulses[0];
            self.ang_impulses = nalgebra::Vector3::new(r.impulses[3], r.impulses[4], r.impulses[5]);
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for PinSlotConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_PIN_SLOT, index_of);
            j.axis1 = v3(&self.axis_v1);
            j.axis3 = v3(&self.axis_w1);
            j.axis2 = v3(&self.axis_w2);
            j.break_force_squared = self.break_force_squared;
            j.break_torque_squared = self.break_torque_squared;
            j.impulses[..3].copy_from_slice(&v3(&self.lin_impulses));
            j.impulses[3..6].copy_from_slice(&v3(&self.ang_impulses));
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.lin_impulses = nalgebra::Vector3::new(r.impulses[0], r.impulses[1], r.impulses[2]);
            self.ang_impulses = nalgebra::Vector3::new(r.impulses[3], r.impulses[4], r.impulses[5]);
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for CylindricalConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_CYLINDRICAL, index_of);
            j.axis1 = v3(&self.axis1);
            j.axis2 = v3(&self.axis2);
            j.break_force_squared = self.break_force_squared;
            j.break_torque_squared = self.break_torque_squared;
            j.impulses[..3].copy_from_slice(&v3(&self.lin_impulses));
            j.impulses[3..6].copy_from_slice(&v3(&self.ang_impulses));
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.lin_impulses = nalgebra::Vector3::new(r.impulses[0], r.impulses[1], r.impulses[2]);
            self.ang_impulses = nalgebra::Vector3::new(r.impulses[3], r.impulses[4], r.impulses[5]);
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for FixedConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_FIXED, index_of);
            j.ref_frame1 = q4(&self.ref_frame1);
            j.ref_frame2 = q4(&self.ref_frame2);
            j.break_force_squared = self.break_force_squared;
            j.break_torque_squared = self.break_torque_squared;
            j.impulses[..3].copy_from_slice(&v3(&self.lin_impulses));
            j.impulses[3..6].copy_from_slice(&v3(&self.ang_impulses));
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.lin_impulses = nalgebra::Vector3::new(r.impulses[0], r.impulses[1], r.impulses[2]);
            self.ang_impulses = nalgebra::Vector3::new(r.impulses[3], r.impulses[4], r.impulses[5]);
            self.broken = r.broken != 0;
        }
    }

    impl<H: BodyHandle> B200Joint<H> for CartesianConstraint<f32, H> {
        fn to_record(&self, index_of: &dyn Fn(H) -> i32) -> sys::nb2_joint {
            let mut j = common!(self, sys::NB2_JOINT_CARTESIAN, index_of);
            j.ref_frame1 = q4(&self.ref_frame1);
            j.ref_frame2 = q4(&self.ref_frame2);
            j.break_torque_squared = self.break_torque_squared;
            j.impulses[3..6].copy_from_slice(&v3(&self.ang_impulses));
            j
        }
        fn store_solver_outputs(&mut self, r: &sys::nb2_joint) {
            self.ang_impulses = nalgebra::Vector3::new(r.impulses[3], r.impulses[4], r.impulses[5]);
            self.broken = r.broken != 0;
        }
    }
}
