"""Reduced-coordinate multibodies on the GPU against the oracle's restatement of
src/object/multibody.rs (oracle/multibody.inc), through the C ABI.  A multibody whose rows only touch static or
kinematic bodies is solved by one thread in the reference's own order, so every mode is compared at 1e-5."""
import numpy as np
import pytest

from nphysics_b200 import abi, scenes

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _pair(sc, params=None):
    from nphysics_b200.solver import Solver
    from oracle import Oracle
    s, o = Solver(0), Oracle()
    for x in (s, o):
        x.set_params(sc.params if params is None else params)
        x.upload_bodies(sc.bodies)
        x.upload_multibodies(sc.multibodies, sc.mb_links)
    return s, o


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def _compare(s, o, what, tol=TOL):
    gl, ol = s.download_multibody_links(), o.download_multibody_links()
    gs, os_ = s.download_body_states(), o.download_body_states()
    errs = {
        "coords": _rel(gl["coords"], ol["coords"]),
        "velocity": _rel(gl["velocity"], ol["velocity"]),
        "impulses": _rel(gl["impulses"], ol["impulses"]),
        "link poses": _rel(gs["position"], os_["position"]),
        "link velocities": _rel(gs["velocity"], os_["velocity"]),
    }
    for k, v in errs.items():
        assert v <= tol, (what, k, v, errs)
    return errs


def test_ragdolls_in_free_flight_match_the_oracle():
    """examples3d/ragdoll3.rs as shipped (FreeJoint torso + five BallJoint members, 21 dofs), no contacts: mass matrix,
    Coriolis terms, LU, accelerations, joint integration.  Teacher-forced?  No: free running, 40 steps."""
    sc = scenes.multibody_ragdolls(3, spin=3.0, colliders=False)
    sc.mb_links["velocity"][1, :3] = [1.0, -2.0, 0.5]   # the head swings
    sc.mb_links["velocity"][9, :3] = [0.0, 0.0, 4.0]    # an arm of the second ragdoll
    s, o = _pair(sc)
    for k in range(40):
        s.step(abi.MODE_COLOURED)
        o.step()
        _compare(s, o, "step %d" % k, tol=5e-5 if k >= 20 else TOL)
    assert s.get_stats()["non_finite"] == 0
    s.close()
    o.close()


@pytest.mark.parametrize("joint", [abi.MBJ_REVOLUTE, abi.MBJ_BALL, abi.MBJ_PRISMATIC])
def test_chains_swinging_from_the_world_match_the_oracle(joint):
    sc = scenes.multibody_chain(joint, links=5, axis=(1, 0, 0) if joint != abi.MBJ_PRISMATIC else (0, 1, 0))
    s, o = _pair(sc)
    for k in range(60):
        s.step(abi.MODE_REFERENCE_ORDER if k % 2 else abi.MODE_COLOURED)  # the multibody path is the same in both modes
        o.step()
        _compare(s, o, "step %d" % k, tol=5e-5)
    s.close()
    o.close()


def test_limits_motors_and_fixed_links_match_the_oracle():
    """Unit-joint rows (unit_joint.rs): a revolute pendulum between two stops, a motor-driven wheel carrying a
    FixedJoint link, a prismatic slider falling onto its lower stop."""
    mb = scenes._ground_only()
    mb.add(-1, abi.MBJ_REVOLUTE, (0.1, 0.1, 0.1), 1.0, parent_shift=(0, 5, 0), body_shift=(0, 0, 0.8), axis=(1, 0, 0),
           flags=abi.MBJ_FLAG_MIN | abi.MBJ_FLAG_MAX, min_pos=-0.3, max_pos=0.2)
    mb.finish()
    w = mb.add(-1, abi.MBJ_REVOLUTE, (0.3, 0.05, 0.3), 1.0, parent_shift=(3, 5, 0), axis=(0, 1, 0),
               flags=abi.MBJ_FLAG_MOTOR, motor_velocity=1.5, motor_max_force=0.05)
    mb.add(w, abi.MBJ_FIXED, (0.05, 0.2, 0.05), 1.0, parent_shift=(0.25, 0.25, 0.0))
    mb.finish()
    mb.add(-1, abi.MBJ_PRISMATIC, (0.1, 0.1, 0.1), 1.0, parent_shift=(-3, 5, 0), axis=(0, 1, 0),
           flags=abi.MBJ_FLAG_MIN | abi.MBJ_FLAG_MAX, min_pos=-0.5, max_pos=0.5)
    mb.finish()
    sc = mb.scene("unit_joints")
    s, o = _pair(sc)
    hit = False
    for k in range(150):
        s.step(abi.MODE_COLOURED)
        o.step()
        e = _compare(s, o, "step %d" % k, tol=5e-5)
        hit = hit or abs(o.download_multibody_links()["impulses"]).max() > 0
    assert hit  # the rows were exercised
    s.close()
    o.close()


def test_ragdolls_landing_on_the_ground_match_the_oracle():
    """Contact rows on multibody links (Multibody::fill_constraint_geometry), impulse cache, position correction with
    update_kinematics after every displacement: ragdolls standing on the ground (feet at the contact margin), then toppling.
    Contacts come from the host producer, re-evaluated every step from the oracle's link poses."""
    sc = scenes.multibody_ragdolls(2, height=2.225 + 0.02, spin=0.0)
    sc.mb_links["velocity"][0, :3] = [0.3, 0.0, 0.1]  # a push, so that they topple
    s, o = _pair(sc)
    gen = scenes.ContactGenerator(sc)
    seen = 0
    for k in range(80):
        m, c = gen.generate(o.download_body_states()["position"])
        seen = max(seen, len(c))
        for x in (s, o):
            x.upload_manifolds(m, c)
        s.step(abi.MODE_COLOURED)
        o.step()
        _compare(s, o, "step %d (%d contacts)" % (k, len(c)), tol=1e-4)
        if len(c):
            assert _rel(s.download_contact_impulses(), o.download_contact_impulses()) <= 1e-4
    s.synchronize()
    assert seen >= 8  # both feet of both ragdolls
    s.close()
    o.close()


def test_composite_joints_as_unit_joints_through_massless_links_match_the_oracle():
    """The reference's multi-dof joints are unit joints side by side (cylindrical_joint.rs:33-90: a PrismaticJoint and a
    RevoluteJoint; universal_joint.rs: two RevoluteJoints; ...).  Through this ABI they are written as what they are: a chain
    of unit joints whose intermediate links carry no mass.  A universal joint (two perpendicular revolutes), a cylindrical
    one (slide + turn about one axis, with a stop) and a planar one (two slides + a turn) swinging from the world."""
    mb = scenes._ground_only()
    a = mb.add(-1, abi.MBJ_REVOLUTE, (0.05, 0.05, 0.05), 0.0, parent_shift=(0, 5, 0), axis=(1, 0, 0), collider=False)
    mb.add(a, abi.MBJ_REVOLUTE, (0.1, 0.3, 0.1), 1.0, body_shift=(0, 0.5, 0.0), axis=(0, 0, 1), coords=[0.4], velocity=[1.0], collider=False)
    mb.finish()
    a = mb.add(-1, abi.MBJ_PRISMATIC, (0.05, 0.05, 0.05), 0.0, parent_shift=(2, 5, 0), axis=(0, 1, 0), collider=False,
               flags=abi.MBJ_FLAG_MIN, min_pos=-0.4)
    mb.add(a, abi.MBJ_REVOLUTE, (0.3, 0.1, 0.1), 1.0, body_shift=(0.2, 0, 0.0), axis=(0, 1, 0), velocity=[2.0], collider=False)
    mb.finish()
    a = mb.add(-1, abi.MBJ_PRISMATIC, (0.05, 0.05, 0.05), 0.0, parent_shift=(4, 5, 0), axis=(1, 0, 0), velocity=[0.5], collider=False)
    b = mb.add(a, abi.MBJ_PRISMATIC, (0.05, 0.05, 0.05), 0.0, axis=(0, 1, 0), collider=False)
    mb.add(b, abi.MBJ_REVOLUTE, (0.3, 0.1, 0.1), 1.0, body_shift=(0.3, 0, 0.0), axis=(0, 0, 1), collider=False)
    mb.finish()
    sc = mb.scene("composite_joints")
    s, o = _pair(sc)
    for k in range(90):
        s.step(abi.MODE_COLOURED)
        o.step()
        _compare(s, o, "step %d" % k, tol=5e-5)
    assert s.get_stats()["non_finite"] == 0
    s.close()
    o.close()


def _run_with_contacts(sc, steps, tol, min_contacts, mode=abi.MODE_COLOURED, tangential=True):
    s, o = _pair(sc)
    gen = scenes.ContactGenerator(sc)
    seen = 0
    for k in range(steps):
        m, c = gen.generate(o.download_body_states()["position"])
        seen = max(seen, len(c))
        for x in (s, o):
            x.upload_manifolds(m, c)
        s.step(mode)
        o.step()
        _compare(s, o, "step %d (%d contacts)" % (k, len(c)), tol=tol)
        if len(c):
            gi, oi = s.download_contact_impulses(), o.download_contact_impulses()
            assert _rel(gi[:, 0], oi[:, 0]) <= tol
            if tangential:
                assert _rel(gi, oi) <= tol
    s.synchronize()
    assert seen >= min_contacts, seen
    s.close()
    o.close()


def test_two_multibodies_stacked_on_the_ground_match_the_oracle():
    """Rows between two different multibodies (both sides Multibody::fill_constraint_geometry; the reference's two-sided
    `unilateral` / `bilateral` classes, solved before the ground classes): a FreeJoint box on a FreeJoint box on the
    ground, plus a third one beside them that only touches the ground.  The two stacked ones are one component."""
    mb = scenes._ground_only((4.0, 0.2, 4.0))
    mb.add(-1, abi.MBJ_FREE, (0.2, 0.1, 0.2), 1.0, coords=[0.0, 0.11, 0.0, 0, 0, 0, 1])
    mb.finish()
    mb.add(-1, abi.MBJ_FREE, (0.1, 0.1, 0.1), 2.0, coords=[0.05, 0.33, -0.03, 0, 0, 0, 1], velocity=[0.2, 0, 0, 0, 0, 0])
    mb.finish()
    mb.add(-1, abi.MBJ_FREE, (0.1, 0.1, 0.1), 1.0, coords=[1.0, 0.11, 0.0, 0, 0, 0, 1])
    mb.finish()
    _run_with_contacts(mb.scene("mb_stack"), 40, 2e-5, 12)


def test_a_multibody_resting_on_its_own_link_matches_the_oracle():
    """Rows between two links of ONE multibody (helper.rs:118-125, the cross terms of `inv_r`): a slider on a vertical
    PrismaticJoint comes down on the FreeJoint box that carries it, which stands on the ground.  The tangential rows
    of that contact are (nearly) redundant with the joint -- J1 + J2 cancels to rounding, r = 1 / (J M^-1 J) is huge -- so
    their impulses are whatever rounding makes them, in both arms, with no effect on the motion: coordinates, velocities
    and normal impulses are compared, the tangential impulses of this scene are not."""
    mb = scenes._ground_only((4.0, 0.2, 4.0))
    root = mb.add(-1, abi.MBJ_FREE, (0.3, 0.1, 0.3), 1.0, coords=[0.0, 0.11, 0.0, 0, 0, 0, 1])
    mb.add(root, abi.MBJ_PRISMATIC, (0.1, 0.1, 0.1), 1.0, parent_shift=(0.0, 0.22, 0.0), axis=(0, 1, 0))
    mb.finish()
    _run_with_contacts(mb.scene("mb_self_contact"), 40, 2e-5, 8, tangential=False)


def test_a_chain_lying_under_a_free_box_matches_the_oracle():
    """A two-link multibody (FreeJoint + BallJoint boxes side by side on the ground) with a single-link multibody on
    top of its second link: contact rows over 9 and 6 generalized coordinates in one component."""
    mb = scenes._ground_only((4.0, 0.2, 4.0))
    root = mb.add(-1, abi.MBJ_FREE, (0.2, 0.1, 0.2), 1.0, coords=[0.0, 0.11, 0.0, 0, 0, 0, 1])
    mb.add(root, abi.MBJ_BALL, (0.2, 0.1, 0.2), 1.0, parent_shift=(0.25, 0.0, 0.0), body_shift=(-0.25, 0.0, 0.0))
    mb.finish()
    mb.add(-1, abi.MBJ_FREE, (0.1, 0.1, 0.1), 1.0, coords=[0.5, 0.33, 0.0, 0, 0, 0, 1])
    mb.finish()
    _run_with_contacts(mb.scene("mb_chain_under_box"), 40, 2e-5, 12)


def test_a_tower_of_multibodies_is_one_component():
    """Five FreeJoint boxes on top of one another (one component, 30 generalized coordinates, rows between consecutive
    members) next to a revolute pendulum with stops that shares nothing with them."""
    mb = scenes._ground_only((4.0, 0.2, 4.0))
    for k in range(5):
        mb.add(-1, abi.MBJ_FREE, (0.2 - 0.02 * k, 0.1, 0.2 - 0.02 * k), 1.0, coords=[0.01 * k, 0.11 + 0.22 * k, 0.0, 0, 0, 0, 1])
        mb.finish()
    mb.add(-1, abi.MBJ_REVOLUTE, (0.1, 0.1, 0.1), 1.0, parent_shift=(2, 3, 0), body_shift=(0, 0, 0.8), axis=(1, 0, 0),
           flags=abi.MBJ_FLAG_MIN | abi.MBJ_FLAG_MAX, min_pos=-0.3, max_pos=0.2, collider=False)
    mb.finish()
    _run_with_contacts(mb.scene("mb_tower"), 60, 5e-5, 20)


def test_multibody_on_a_kinematic_platform_and_the_frictionless_model_match_the_oracle():
    """A FreeJoint box riding a KINEMATIC platform that moves sideways and up (the platform's velocity enters the rows'
    right-hand sides, rigid_body.rs:693-698), under the pyramid model and under SignoriniModel (frictionless, rows for
    active contacts only)."""
    for model in (abi.CONTACT_SIGNORINI_COULOMB_PYRAMID, abi.CONTACT_SIGNORINI):
        mb = scenes._ground_only((4.0, 0.2, 4.0))
        mb.add(-1, abi.MBJ_FREE, (0.1, 0.1, 0.1), 1.0, coords=[0.0, 0.11, 0.0, 0, 0, 0, 1])
        mb.finish()
        sc = mb.scene("mb_platform")
        sc.bodies["status"][0] = abi.BODY_KINEMATIC
        sc.bodies["velocity"][0, :3] = [0.3, 0.2, 0.0]
        s, o = _pair(sc)
        for x in (s, o):
            x.set_contact_model(model)
        gen = scenes.ContactGenerator(sc)
        for k in range(40):
            m, c = gen.generate(o.download_body_states()["position"])
            for x in (s, o):
                x.upload_manifolds(m, c)
            s.step(abi.MODE_COLOURED)
            o.step()
            _compare(s, o, "model %d step %d" % (model, k), tol=2e-5)
            assert _rel(s.download_contact_impulses()[:, 0], o.download_contact_impulses()[:, 0]) <= 2e-5
        states = o.download_body_states()
        assert states["position"][0, 0] > 0.15 and states["position"][1, 1] > 0.2  # the platform carried the box up
        if model == abi.CONTACT_SIGNORINI_COULOMB_PYRAMID:
            assert states["position"][1, 0] > 0.1  # friction drags it along; the frictionless model leaves it behind
        else:
            assert abs(states["position"][1, 0]) < 1e-3
        s.close()
        o.close()


def test_a_multibody_touching_a_dynamic_body_is_reported():
    mb = scenes._ground_only()
    mb.add(-1, abi.MBJ_FREE, (0.1, 0.1, 0.1), 1.0, coords=[0.0, 0.11, 0.0, 0, 0, 0, 1])
    mb.finish()
    sc = mb.scene("coupled")
    bodies = np.concatenate([sc.bodies, abi.new_bodies(1)])
    bodies["position"][2, :3] = [0.0, 0.33, 0.0]
    bodies["mass"][2] = 1.0
    bodies["local_inertia"][2] = np.eye(3).reshape(9)
    from nphysics_b200.solver import Solver
    s = Solver(0)
    s.upload_bodies(bodies)
    s.upload_multibodies(sc.multibodies, sc.mb_links)
    m = np.zeros(1, dtype=abi.manifold_dtype)
    m["body1"], m["body2"], m["num_contacts"] = 1, 2, 1
    m["coll1_wrt_body"][0, 6] = m["coll2_wrt_body"][0, 6] = 1.0
    c = np.zeros(1, dtype=abi.contact_dtype)
    c["normal"][0] = [0, 1, 0]
    c["key"] = 1
    c["geom1"], c["geom2"] = abi.GEOM_PLANE, abi.GEOM_POINT
    s.upload_manifolds(m, c)
    s.step(abi.MODE_COLOURED)
    with pytest.raises(RuntimeError):
        s.synchronize()
    s.close()
