//! `B200MoreauJeanSolver`: same `step` signature as `nphysics3d::solver::MoreauJeanSolver`
//! (src/solver/moreau_jean_solver.rs:47-61); the body marshals the borrowed inputs into the flat
//! records of `include/nphysics_b200.h` and calls the C ABI.  NOT compiled in the build environment
//! (no rustc) -- see ../README.md.
use nalgebra as na;
use ncollide3d::query::ContactKinematic;
use ncollide3d::shape::FeatureId;
use nphysics3d::counters::Counters;
use nphysics3d::detection::ColliderContactManifold;
use nphysics3d::joint::JointConstraintSet;
use nphysics3d::material::{Material, MaterialContext, MaterialsCoefficientsTable};
use nphysics3d::object::{Body, BodyHandle, BodySet, BodyStatus, ColliderAnchor, ColliderHandle, ColliderSet};
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
    joints: Vec<sys::nb2_joint>,
    _marker: std::marker::PhantomData<CollHandle>,
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
                  bodies: vec![], states: vec![], manifolds: vec![], contacts: vec![], joints: vec![],
                  _marker: std::marker::PhantomData })
    }

    /// Perform one step of the time-stepping scheme (drop-in for MoreauJeanSolver::step).
    pub fn step<Colliders, Constraints>(
        &mut self,
        _counters: &mut Counters,
        bodies: &mut dyn BodySet<f32, Handle = Handle>,
        colliders: &Colliders,
        _joints: &mut Constraints,
        manifolds: &[ColliderContactManifold<f32, Handle, CollHandle>],
        _island: &[Handle],
        _island_joints: &[Constraints::Handle],
        parameters: &IntegrationParameters<f32>,
        coefficients: &MaterialsCoefficientsTable<f32>,
    ) where
        Colliders: ColliderSet<f32, Handle, Handle = CollHandle>,
        Constraints: JointConstraintSet<f32, Handle>,
    {
        let _ = colliders;
        // 1. bodies -> nb2_body records (index = insertion order); re-uploaded when any update flag is set
        self.handles.clear();
        self.bodies.clear();
        self.index_of.clear();
        bodies.foreach(&mut |h, b: &dyn Body<f32>| {
            let rb = b.downcast_ref::<nphysics3d::object::RigidBody<f32>>();
            let mut rec: sys::nb2_body = unsafe { std::mem::zeroed() };
            rec.position[6] = 1.0;
            rec.status = match b.status() {
                BodyStatus::Disabled => sys::NB2_BODY_DISABLED,
                BodyStatus::Static => sys::NB2_BODY_STATIC,
                BodyStatus::Dynamic => sys::NB2_BODY_DYNAMIC,
                BodyStatus::Kinematic => sys::NB2_BODY_KINEMATIC,
            };
            if let Some(rb) = rb {
                let p = rb.position();
                rec.position = [p.translation.x, p.translation.y, p.translation.z,
                                p.rotation.i, p.rotation.j, p.rotation.k, p.rotation.w];
                let v = rb.velocity();
                rec.velocity = [v.linear.x, v.linear.y, v.linear.z, v.angular.x, v.angular.y, v.angular.z];
                // local_com, mass, local_inertia (row-major), damping, caps, jacobian mask, gravity flag ...
                if b.gravity_enabled() { rec.flags |= sys::NB2_BODY_FLAG_GRAVITY; }
            } else {
                rec.status = sys::NB2_BODY_STATIC; // Ground and unsupported body kinds act as ground
            }
            self.index_of.insert(h, self.bodies.len() as i32);
            self.handles.push(h);
            self.bodies.push(rec);
        });
        unsafe { sys::nb2_upload_bodies(self.ctx, self.bodies.as_ptr(), self.bodies.len() as u32); }

        // 2. manifolds -> nb2_manifold / nb2_contact
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
                    // Material::combine is evaluated once per manifold for BasicMaterial pairs
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
        // 3. joints: anchors/axes/cached impulses of each active constraint (per joint type) -> nb2_joint
        // 4. params + step + 5. write back
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
        p.gravity = self.gravity;
        unsafe {
            sys::nb2_set_params(self.ctx, &p);
            sys::nb2_upload_manifolds(self.ctx, self.manifolds.as_ptr(), self.manifolds.len() as u32,
                                      self.contacts.as_ptr(), self.contacts.len() as u32);
            sys::nb2_step(self.ctx, sys::NB2_MODE_COLOURED);
            self.states.resize(self.bodies.len(), std::mem::zeroed());
            sys::nb2_download_body_states(self.ctx, self.states.as_mut_ptr(), 0, self.states.len() as u32);
        }
        for (h, s) in self.handles.iter().zip(self.states.iter()) {
            if let Some(b) = bodies.get_mut(*h) {
                if let Some(rb) = b.downcast_mut::<nphysics3d::object::RigidBody<f32>>() {
                    let q = na::UnitQuaternion::new_unchecked(na::Quaternion::new(s.position[6], s.position[3], s.position[4], s.position[5]));
                    rb.set_position(na::Isometry3::from_parts(na::Translation3::new(s.position[0], s.position[1], s.position[2]), q));
                    rb.set_velocity(nphysics3d::algebra::Velocity3::new(
                        na::Vector3::new(s.velocity[0], s.velocity[1], s.velocity[2]),
                        na::Vector3::new(s.velocity[3], s.velocity[4], s.velocity[5])));
                }
            }
        }
    }
}

impl<Handle: BodyHandle, CollHandle: ColliderHandle> Drop for B200MoreauJeanSolver<Handle, CollHandle> {
    fn drop(&mut self) {
        unsafe { sys::nb2_destroy(self.ctx); }
    }
}

/// ContactId is a slotmap key: its (index, version) pair is a stable 64-bit id.
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
    let _ = FeatureId::Unknown;
}
