//! `B200MoreauJeanSolver`: same `step` signature as `nphysics3d::solver::MoreauJeanSolver`
//! (src/solver/moreau_jean_solver.rs:47-61); the body marshals the borrowed inputs into the flat
//! records of `include/nphysics_b200.h` and calls the C ABI.  NOT compiled in the build environment
//! (no rustc) -- see ../README.md.  Everything the C ABI needs is marshalled here:
//!
//!   bodies     every field of `nb2_body` (mass properties, damping, velocity caps, kinematic dofs,
//!              external forces, gravity flag); the whole set is re-uploaded only when a body's
//!              `update_status()` reports more than a pose / velocity change or the set itself changed --
//!              otherwise only the touched poses and velocities go up (`nb2_upload_body_states`)
//!   manifolds  `nb2_manifold` + `nb2_contact` from ncollide's manifolds; when the contact set kept its
//!              shape since the last step only the 40-byte per-step part travels (`nb2_update_contacts`)
//!   joints     through `joints::B200Joint` (ten constraint types, cached impulses and `broken` both ways)
//!   results    poses and velocities written back into the bodies, impulses / broken flags into the joints
pub mod joints;

use joints::B200Joint;
use nalgebra as na;
use ncollide3d::query::ContactKinematic;
use nphysics3d::counters::Counters;
use nphysics3d::detection::ColliderContactManifold;
use nphysics3d::joint::{BallJoint, FixedJoint, FreeJoint, Joint, JointConstraintSet, JointMotor, PrismaticJoint, RevoluteJoint};
use nphysics3d::material::{Material, MaterialContext, MaterialsCoefficientsTable};
use nphysics3d::object::{Body, BodyHandle, BodyPart, BodySet, BodyStatus, BodyUpdateStatus, ColliderAnchor, ColliderHandle,
                         ColliderSet, RigidBody};
use nphysics3d::object::Multibody;
use nphysics3d::solver::IntegrationParameters;
use nphysics_b200_sys as sys;
use std::collections::HashMap;

pub struct B200MoreauJeanSolver<Handle: BodyHandle, CollHandle: ColliderHandle> {
    ctx: *mut sys::nb2_context,
    gravity: [f32; 3],
    index_of: HashMap<Handle, i32>,
    handles: Vec<Handle>,
    bodies: Vec<sys::nb2_body>,
    states: Vec<sys::nb2_body_state>,
    manifolds: Vec<sys::nb2_manifold>,
    contacts: Vec<sys::nb2_contact>,
    updates: Vec<sys::nb2_contact_update>,
    /// (body1, body2, first_contact, num_contacts) + contact ids of the last full upload
    uploaded_shape: Vec<(i32, i32, u32, u32)>,
    uploaded_ids: Vec<u64>,
    joints: Vec<sys::nb2_joint>,
    mode: i32,
    _marker: std::marker::PhantomData<CollHandle>,
}

fn check(ctx: *mut sys::nb2_context, rc: i32) {
    // the reference's step has no error channel: invariants panic (moreau_jean_solver.rs:147-150)
    if rc != sys::NB2_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::nb2_last_error(ctx)) };
        panic!("nphysics-b200: {}", msg.to_string_lossy());
    }
}

impl<Handle: BodyHandle, CollHandle: ColliderHandle> B200MoreauJeanSolver<Handle, CollHandle> {
    pub fn new(device: i32, gravity: na::Vector3<f32>) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::nb2_create(device, std::ptr::null_mut(), &mut ctx) };
        if rc != sys::NB2_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(sys::nb2_last_error(std::ptr::null())) };
            return Err(msg.to_string_lossy().into_owned());
        }
        Ok(Self { ctx, gravity: [gravity.x, gravity.y, gravity.z], index_of: HashMap::new(), handles: vec![],
                  bodies: vec![], states: vec![], manifolds: vec![], contacts: vec![], updates: vec![],
                  uploaded_shape: vec![], uploaded_ids: vec![], joints: vec![], mode: sys::NB2_MODE_COLOURED,
                  _marker: std::marker::PhantomData })
    }

    /// `NB2_MODE_REFERENCE_ORDER` replays the reference's sequential sweep (verification);
    /// `NB2_MODE_COLOURED` (default) is the production mode.
    pub fn set_mode(&mut self, mode: i32) { self.mode = mode; }

    fn body_record(b: &dyn Body<f32>) -> sys::nb2_body {
        let mut rec: sys::nb2_body = unsafe { std::mem::zeroed() };
        rec.position[6] = 1.0;
        rec.jacobian_mask = [1.0; 6];
        rec.max_linear_velocity = f32::MAX;
        rec.max_angular_velocity = f32::MAX;
        rec.status = match b.status() {
            BodyStatus::Disabled => sys::NB2_BODY_DISABLED,
            BodyStatus::Static => sys::NB2_BODY_STATIC,
            BodyStatus::Dynamic => sys::NB2_BODY_DYNAMIC,
            BodyStatus::Kinematic => sys::NB2_BODY_KINEMATIC,
        };
        if let Some(rb) = b.downcast_ref::<RigidBody<f32>>() {
            let p = rb.position();
            rec.position = [p.translation.x, p.translation.y, p.translation.z,
                            p.rotation.i, p.rotation.j, p.rotation.k, p.rotation.w];
            let v = rb.velocity();
            rec.velocity = [v.linear.x, v.linear.y, v.linear.z, v.angular.x, v.angular.y, v.angular.z];
            let com = rb.local_center_of_mass();           // BodyPart (rigid_body.rs:928)
            rec.local_com = [com.x, com.y, com.z];
            let inertia = rb.local_inertia();              // Inertia3 { linear, angular } (rigid_body.rs:913)
            rec.mass = inertia.linear;
            for r in 0..3 { for c in 0..3 { rec.local_inertia[3 * r + c] = inertia.angular[(r, c)]; } }
            rec.linear_damping = rb.linear_damping();
            rec.angular_damping = rb.angular_damping();
            rec.max_linear_velocity = rb.max_linear_velocity();
            rec.max_angular_velocity = rb.max_angular_velocity();
            // jacobian_mask: 0 on kinematic dofs (rigid_body.rs:105-121)
            let (kt, kr) = (rb.kinematic_translations(), rb.kinematic_rotations());
            for k in 0..3 {
                rec.jacobian_mask[k] = if kt[k] { 0.0 } else { 1.0 };
                rec.jacobian_mask[3 + k] = if kr[k] { 0.0 } else { 1.0 };
            }
            // external forces: `RigidBody::external_forces` is private (rigid_body.rs:36); the accessor added by
            // the in-crate patch returns the accumulated Force3
            let f = rb.b200_external_forces();
            rec.external_forces = [f.linear.x, f.linear.y, f.linear.z, f.angular.x, f.angular.y, f.angular.z];
            if b.gravity_enabled() { rec.flags |= sys::NB2_BODY_FLAG_GRAVITY; }
        } else {
            rec.status = sys::NB2_BODY_STATIC; // Ground (ground.rs) and body kinds outside SURVEY section 8 act as ground
        }
        rec
    }

    /// Perform one step of the time-stepping scheme (drop-in for MoreauJeanSolver::step).
    pub fn step<Colliders, Constraints>(
        &mut self,
        _counters: &mut Counters,
        bodies: &mut dyn BodySet<f32, Handle = Handle>,
        _colliders: &Colliders,
        joints: &mut Constraints,
        manifolds: &[ColliderContactManifold<f32, Handle, CollHandle>],
        island: &[Handle],
        island_joints: &[Constraints::Handle],
        parameters: &IntegrationParameters<f32>,
        coefficients: &MaterialsCoefficientsTable<f32>,
    ) where
        Colliders: ColliderSet<f32, Handle, Handle = CollHandle>,
        Constraints: JointConstraintSet<f32, Handle>,
        Constraints::JointConstraint: B200Joint<Handle>,
    {
        let ctx = self.ctx;
        // ---- 1. bodies.  The uploaded set is every body of the BodySet in iteration order (bodies outside `island`
        // -- sleeping, static -- must still be addressable by manifolds); a change of the set or of anything but
        // pose / velocity re-uploads it, otherwise only edited poses / velocities travel.
        let mut same_set = true;
        let mut heavy_change = false;
        let mut k = 0usize;
        bodies.foreach(&mut |h, b: &dyn Body<f32>| {
            if k >= self.handles.len() || self.handles[k] != h { same_set = false; }
            let st = b.update_status();
            if st.inertia_changed() || st.local_inertia_changed() || st.local_com_changed() || st.damping_changed() || st.status_changed() {
                heavy_change = true;
            }
            k += 1;
        });
        if k != self.handles.len() { same_set = false; }
        if !same_set || heavy_change {
            self.handles.clear();
            self.bodies.clear();
            self.index_of.clear();
            bodies.foreach(&mut |h, b: &dyn Body<f32>| {
                self.index_of.insert(h, self.bodies.len() as i32);
                self.handles.push(h);
                self.bodies.push(Self::body_record(b));
            });
            check(ctx, unsafe { sys::nb2_upload_bodies(ctx, self.bodies.as_ptr(), self.bodies.len() as u32) });
            self.uploaded_shape.clear();
        } else {
            // dirty ranges: consecutive bodies whose pose or velocity was edited by the user since the last step
            let mut i = 0usize;
            let mut run: Option<(usize, Vec<sys::nb2_body_state>)> = None;
            let mut flush = |run: &mut Option<(usize, Vec<sys::nb2_body_state>)>| {
                if let Some((first, st)) = run.take() {
                    check(ctx, unsafe { sys::nb2_upload_body_states(ctx, st.as_ptr(), first as u32, st.len() as u32) });
                    check(ctx, unsafe { sys::nb2_synchronize(ctx) });  // `st` is dropped on return
                }
            };
            bodies.foreach(&mut |_h, b: &dyn Body<f32>| {
                let st = b.update_status();
                if st.position_changed() || st.velocity_changed() {
                    let rec = Self::body_record(b);
                    let s = sys::nb2_body_state { position: rec.position, velocity: rec.velocity };
                    match run.as_mut() {
                        Some((_, v)) => v.push(s),
                        None => run = Some((i, vec![s])),
                    }
                } else {
                    flush(&mut run);
                }
                i += 1;
            });
            flush(&mut run);
        }
        let _ = island; // the device filters by effective status exactly as `island` does (mechanical_world.rs:287-313)

        // ---- 2. manifolds -> nb2_manifold / nb2_contact
        self.manifolds.clear();
        self.contacts.clear();
        for m in manifolds {
            let mut rec: sys::nb2_manifold = unsafe { std::mem::zeroed() };
            rec.body1 = self.index_of[&m.body1()];
            rec.body2 = self.index_of[&m.body2()];
            rec.first_contact = self.contacts.len() as u32;
            rec.margin1 = m.collider1.margin();
            rec.margin2 = m.collider2.margin();
            rec.coll1_wrt_body[6] = 1.0;
            rec.coll2_wrt_body[6] = 1.0;
            if let ColliderAnchor::OnBodyPart { position_wrt_body_part: p, .. } = m.collider1.anchor() {
                rec.coll1_wrt_body = [p.translation.x, p.translation.y, p.translation.z, p.rotation.i, p.rotation.j, p.rotation.k, p.rotation.w];
            }
            if let ColliderAnchor::OnBodyPart { position_wrt_body_part: p, .. } = m.collider2.anchor() {
                rec.coll2_wrt_body = [p.translation.x, p.translation.y, p.translation.z, p.rotation.i, p.rotation.j, p.rotation.k, p.rotation.w];
            }
            for c in m.contacts() {
                if rec.num_contacts == 0 {
                    // Material::combine is evaluated once per manifold for BasicMaterial pairs (material.rs:134-177)
                    let ctx1 = MaterialContext::new(m.collider1.shape(), m.collider1.position(), c, true);
                    let ctx2 = MaterialContext::new(m.collider2.shape(), m.collider2.position(), c, false);
                    let props = <dyn Material<f32>>::combine(coefficients, m.collider1.material(), ctx1, m.collider2.material(), ctx2);
                    rec.friction = props.friction.0;
                    rec.restitution = props.restitution.0;
                    rec.surface_velocity = [props.surface_velocity.x, props.surface_velocity.y, props.surface_velocity.z];
                }
                let mut cr: sys::nb2_contact = unsafe { std::mem::zeroed() };
                cr.world1 = [c.contact.world1.x, c.contact.world1.y, c.contact.world1.z];
                cr.world2 = [c.contact.world2.x, c.contact.world2.y, c.contact.world2.z];
                cr.normal = [c.contact.normal.x, c.contact.normal.y, c.contact.normal.z];
                cr.depth = c.contact.depth;
                cr.key = contact_key(&c.id);
                fill_kinematic(&mut cr, &c.kinematic);
                self.contacts.push(cr);
                rec.num_contacts += 1;
            }
            self.manifolds.push(rec);
        }
        // the same contacts as last step (ids, order, manifolds): only world1 / world2 / normal / depth go up
        let same_shape = self.uploaded_shape.len() == self.manifolds.len()
            && self.uploaded_ids.len() == self.contacts.len()
            && self.manifolds.iter().zip(self.uploaded_shape.iter())
                   .all(|(m, s)| (m.body1, m.body2, m.first_contact, m.num_contacts) == *s)
            && self.contacts.iter().zip(self.uploaded_ids.iter()).all(|(c, id)| c.key == *id);
        if same_shape {
            self.updates.clear();
            self.updates.extend(self.contacts.iter().map(|c| sys::nb2_contact_update {
                world1: c.world1, world2: c.world2, normal: c.normal, depth: c.depth }));
            check(ctx, unsafe { sys::nb2_update_contacts(ctx, self.updates.as_ptr(), self.updates.len() as u32) });
        } else {
            check(ctx, unsafe { sys::nb2_upload_manifolds(ctx, self.manifolds.as_ptr(), self.manifolds.len() as u32,
                                                          self.contacts.as_ptr(), self.contacts.len() as u32) });
            self.uploaded_shape = self.manifolds.iter().map(|m| (m.body1, m.body2, m.first_contact, m.num_contacts)).collect();
            self.uploaded_ids = self.contacts.iter().map(|c| c.key).collect();
        }

        // ---- 3. joints: the active constraints in island_joints order (mechanical_world.rs:274-279)
        self.joints.clear();
        {
            let index_of = &self.index_of;
            let lookup = |h: Handle| -> i32 { index_of[&h] };
            for jh in island_joints {
                if let Some(j) = joints.get(*jh) {
                    self.joints.push(j.to_record(&lookup));
                }
            }
        }
        check(ctx, unsafe { sys::nb2_upload_joints(ctx, self.joints.as_ptr(), self.joints.len() as u32) });

        // ---- 4. parameters + step
        let mut p: sys::nb2_params = unsafe { std::mem::zeroed() };
        unsafe { sys::nb2_default_params(&mut p); }
        p.dt = parameters.dt();
        p.erp = parameters.erp;
        p.warmstart_coeff = parameters.warmstart_coeff;
        p.restitution_velocity_threshold = parameters.restitution_velocity_threshold;
        p.allowed_linear_error = parameters.allowed_linear_error;
        p.allowed_angular_error = parameters.allowed_angular_error;
        p.max_linear_correction = parameters.max_linear_correction;
        p.max_angular_correction = parameters.max_angular_correction;
        p.max_stabilization_multiplier = parameters.max_stabilization_multiplier;
        p.max_velocity_iterations = parameters.max_velocity_iterations as u32;
        p.max_position_iterations = parameters.max_position_iterations as u32;
        p.max_ccd_position_iterations = parameters.max_ccd_position_iterations as u32;
        p.max_ccd_substeps = parameters.max_ccd_substeps as u32;
        p.gravity = self.gravity;
        check(ctx, unsafe { sys::nb2_set_params(ctx, &p) });
        check(ctx, unsafe { sys::nb2_step(ctx, self.mode) });

        // ---- 5. write back: poses and velocities into the bodies, cached impulses and broken flags into the joints
        self.states.resize(self.bodies.len(), unsafe { std::mem::zeroed() });
        check(ctx, unsafe { sys::nb2_download_body_states(ctx, self.states.as_mut_ptr(), 0, self.states.len() as u32) });
        for (h, s) in self.handles.iter().zip(self.states.iter()) {
            if let Some(b) = bodies.get_mut(*h) {
                if let Some(rb) = b.downcast_mut::<RigidBody<f32>>() {
                    if rb.status() == BodyStatus::Dynamic || rb.status() == BodyStatus::Kinematic {
                        let q = na::UnitQuaternion::new_unchecked(na::Quaternion::new(s.position[6], s.position[3], s.position[4], s.position[5]));
                        rb.set_position(na::Isometry3::from_parts(na::Translation3::new(s.position[0], s.position[1], s.position[2]), q));
                        rb.set_velocity(nphysics3d::algebra::Velocity3::new(
                            na::Vector3::new(s.velocity[0], s.velocity[1], s.velocity[2]),
                            na::Vector3::new(s.velocity[3], s.velocity[4], s.velocity[5])));
                    }
                }
                // what the device now holds is what the host holds: nothing is dirty (MechanicalWorld::step clears the
                // flags at the same place, mechanical_world.rs:340-346)
                b.clear_update_flags();
            }
        }
        if !self.joints.is_empty() {
            check(ctx, unsafe { sys::nb2_download_joints(ctx, self.joints.as_mut_ptr(), self.joints.len() as u32) });
            let mut k = 0usize;
            for jh in island_joints {
                if let Some(j) = joints.get_mut(*jh) {
                    j.store_solver_outputs(&self.joints[k]);
                    k += 1;
                }
            }
        }
    }

    /// cuboid piles can leave their narrow phase on the device as well.  Call once with the
    /// colliders (`nb2_collider` records, `body` = index in BodySet iteration order), then `step_with_device_contacts`
    /// instead of `step`: no contact data crosses PCIe at all.
    pub fn use_device_contacts(&mut self, colliders: &[sys::nb2_collider], linear_prediction: f32) -> u32 {
        let mut n_pairs = 0u32;
        check(self.ctx, unsafe { sys::nb2_upload_colliders(self.ctx, colliders.as_ptr(), colliders.len() as u32) });
        check(self.ctx, unsafe { sys::nb2_detect_pairs(self.ctx, linear_prediction, -1.0, 0, &mut n_pairs) });
        n_pairs
    }

    pub fn step_with_device_contacts(&mut self) {
        check(self.ctx, unsafe { sys::nb2_generate_manifolds(self.ctx) });
        check(self.ctx, unsafe { sys::nb2_step(self.ctx, self.mode) });
    }
}

impl<Handle: BodyHandle, CollHandle: ColliderHandle> Drop for B200MoreauJeanSolver<Handle, CollHandle> {
    fn drop(&mut self) {
        unsafe { sys::nb2_destroy(self.ctx); }
    }
}

/// ContactId is a slotmap key: its (index, version) pair is a stable 64-bit id.
// ---------------------------------------------------------------------------------------------------------------
// Reduced-coordinate multibodies.  A `Multibody` of the body set becomes one `nb2_multibody` and one
// `nb2_mb_link` per link; every link also takes a `nb2_body` record of status NB2_BODY_MULTIBODY_LINK (the record a
// `BodyPartHandle(handle, i)` resolves to for colliders and manifolds).  The reference keeps the joint coordinates
// private to each `Joint` implementation (free_joint.rs:13, ball_joint.rs:14, revolute_joint.rs:24, ...): the accessors
// used below -- `FreeJoint::position()`, `BallJoint::rotation()`, `FixedJoint::body_to_parent()` and their setters --
// are the (one-line) additions the integration needs next to the existing `RevoluteJoint::angle()` /
// `PrismaticJoint::offset()`; the generalized velocities, damping and mass properties are public already
// (`Multibody::generalized_velocity`, `damping`, `MultibodyLink` + `BodyPart::local_inertia / local_center_of_mass`).
pub fn marshal_multibody(mb: &Multibody<f32>, mb_index: i32, first_body: i32,
                         links: &mut Vec<sys::nb2_mb_link>, parts: &mut Vec<sys::nb2_body>) -> sys::nb2_multibody {
    let first_link = links.len() as u32;
    let vels = mb.generalized_velocity();
    let damping = mb.damping();
    let mut dof = 0usize;
    for (i, link) in mb.links().enumerate() {
        let mut l: sys::nb2_mb_link = unsafe { std::mem::zeroed() };
        l.multibody = mb_index;
        l.parent = link.parent_id().map(|p| p as i32).unwrap_or(-1);
        l.body = first_body + i as i32;
        l.parent_shift = [link.parent_shift().x, link.parent_shift().y, link.parent_shift().z];
        l.body_shift = [link.body_shift().x, link.body_shift().y, link.body_shift().z];
        l.axis = [1.0, 0.0, 0.0];
        l.motor_max_velocity = f32::MAX;
        l.motor_max_force = f32::MAX;
        let joint = link.joint();
        let ndofs = joint.ndofs();
        let iso = |p: &na::Isometry3<f32>| [p.translation.x, p.translation.y, p.translation.z,
                                            p.rotation.i, p.rotation.j, p.rotation.k, p.rotation.w];
        let unit = |l: &mut sys::nb2_mb_link, min: Option<f32>, max: Option<f32>, motor: &JointMotor<f32, f32>| {
            if let Some(v) = min { l.flags |= sys::NB2_MBJ_FLAG_MIN; l.min_pos = v; }
            if let Some(v) = max { l.flags |= sys::NB2_MBJ_FLAG_MAX; l.max_pos = v; }
            if motor.enabled { l.flags |= sys::NB2_MBJ_FLAG_MOTOR; }
            l.motor_velocity = motor.desired_velocity;
            l.motor_max_velocity = motor.max_velocity;
            l.motor_max_force = motor.max_force;
        };
        if let Some(j) = joint.downcast_ref::<FreeJoint<f32>>() {
            l.joint_type = sys::NB2_MBJ_FREE;
            l.coords = iso(j.position());
        } else if let Some(j) = joint.downcast_ref::<BallJoint<f32>>() {
            l.joint_type = sys::NB2_MBJ_BALL;
            let r = j.rotation();
            l.coords[..4].copy_from_slice(&[r.i, r.j, r.k, r.w]);
        } else if let Some(j) = joint.downcast_ref::<RevoluteJoint<f32>>() {
            l.joint_type = sys::NB2_MBJ_REVOLUTE;
            l.axis = [j.axis().x, j.axis().y, j.axis().z];
            l.coords[0] = j.angle();
            unit(&mut l, j.min_angle(), j.max_angle(), j.motor());
        } else if let Some(j) = joint.downcast_ref::<PrismaticJoint<f32>>() {
            l.joint_type = sys::NB2_MBJ_PRISMATIC;
            l.axis = [j.axis().x, j.axis().y, j.axis().z];
            l.coords[0] = j.offset();
            unit(&mut l, j.min_offset(), j.max_offset(), j.motor());
        } else if let Some(j) = joint.downcast_ref::<FixedJoint<f32>>() {
            l.joint_type = sys::NB2_MBJ_FIXED;
            l.coords = iso(j.body_to_parent());
        } else {
            panic!("nphysics-b200: this reduced-coordinate joint has no device form (DESIGN.md section 8b)");
        }
        for d in 0..ndofs {
            l.velocity[d] = vels[dof + d];
            l.damping[d] = damping[dof + d];
        }
        if ndofs == 1 {  // cached impulses of the motor / min / max rows (unit_joint.rs:77, 113, 153)
            let imp = mb.impulses();
            l.impulses.copy_from_slice(&imp[3 * dof..3 * dof + 3]);
        }
        dof += ndofs;
        links.push(l);
        // the link as a body record: mass properties in, pose and velocity out
        let mut b: sys::nb2_body = unsafe { std::mem::zeroed() };
        b.status = sys::NB2_BODY_MULTIBODY_LINK;
        b.position[6] = 1.0;
        b.mass = link.local_inertia().linear;
        let li = link.local_inertia().angular;
        b.local_inertia = [li.m11, li.m12, li.m13, li.m21, li.m22, li.m23, li.m31, li.m32, li.m33];
        let lc = link.local_center_of_mass();
        b.local_com = [lc.x, lc.y, lc.z];
        b.jacobian_mask = [1.0; 6];
        b.max_linear_velocity = f32::MAX;
        b.max_angular_velocity = f32::MAX;
        parts.push(b);
    }
    sys::nb2_multibody {
        first_link,
        n_links: links.len() as u32 - first_link,
        flags: if mb.gravity_enabled() { sys::NB2_BODY_FLAG_GRAVITY } else { 0 },
        reserved: 0,
    }
}

/// After the step: joint coordinates, generalized velocities and cached impulses back into the Multibody
/// (`links` = what nb2_download_multibody_links returned for this multibody).
pub fn unmarshal_multibody(mb: &mut Multibody<f32>, links: &[sys::nb2_mb_link]) {
    let mut dof = 0usize;
    for (i, l) in links.iter().enumerate() {
        let ndofs = {
            let joint = mb.link_mut(i).unwrap().joint_mut();
            let q = |c: &[f32]| na::UnitQuaternion::new_unchecked(na::Quaternion::new(c[3], c[0], c[1], c[2]));
            if let Some(j) = joint.downcast_mut::<FreeJoint<f32>>() {
                j.set_position(na::Isometry3::from_parts(na::Translation3::new(l.coords[0], l.coords[1], l.coords[2]), q(&l.coords[3..7])));
            } else if let Some(j) = joint.downcast_mut::<BallJoint<f32>>() {
                j.set_rotation(q(&l.coords[0..4]));
            } else if let Some(j) = joint.downcast_mut::<RevoluteJoint<f32>>() {
                j.set_angle(l.coords[0]);
            } else if let Some(j) = joint.downcast_mut::<PrismaticJoint<f32>>() {
                j.set_offset(l.coords[0]);
            }
            joint.ndofs()
        };
        {
            let mut v = mb.generalized_velocity_mut();
            for d in 0..ndofs { v[dof + d] = l.velocity[d]; }
        }
        dof += ndofs;
    }
    mb.update_kinematics();  // link poses from the new coordinates, as MechanicalWorld::step does at :343-346
}

fn contact_key(id: &ncollide3d::query::ContactId) -> u64 {
    use slotmap::Key;
    id.data().as_ffi()
}

fn fill_kinematic(cr: &mut sys::nb2_contact, k: &ContactKinematic<f32>) {
    use ncollide3d::query::NeighborhoodGeometry as G;
    let (a1, a2) = (k.approx1(), k.approx2());
    cr.local1 = [a1.point.x, a1.point.y, a1.point.z];
    cr.local2 = [a2.point.x, a2.point.y, a2.point.z];
    cr.dilation1 = k.dilation1();
    cr.dilation2 = k.dilation2();
    let tag = |g: &G<f32>, dir: &mut [f32; 3]| -> u8 {
        match g {
            G::Point => sys::NB2_GEOM_POINT,
            G::Line(d) => { *dir = [d.x, d.y, d.z]; sys::NB2_GEOM_LINE }
            G::Plane(n) => { *dir = [n.x, n.y, n.z]; sys::NB2_GEOM_PLANE }
        }
    };
    cr.geom1 = tag(&a1.geometry, &mut cr.dir1);
    cr.geom2 = tag(&a2.geometry, &mut cr.dir2);
}
