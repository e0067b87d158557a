"""What pins the oracle's reduced-coordinate multibody (oracle/multibody.inc): closed forms and
conservation laws of the algorithm of src/object/multibody.rs.  PARITY UNPINNED like the rest of the oracle: the
reference has no test for this path and cannot be compiled here."""
import numpy as np
import pytest

from nphysics_b200 import abi, scenes
from oracle import Oracle


def _run(sc, steps, params=None, f64=False, manifolds=None):
    o = Oracle(f64=f64)
    o.set_params(sc.params if params is None else params)
    o.upload_bodies(sc.bodies)
    o.upload_multibodies(sc.multibodies, sc.mb_links)
    if manifolds is not None:
        o.upload_manifolds(*manifolds)
    out = []
    for _ in range(steps):
        o.step()
        out.append((o.download_body_states().copy(), o.download_multibody_links().copy()))
    o.close()
    return out


def _free_box(vel, spin, gravity=True):
    mb = scenes._ground_only()
    mb.add(-1, abi.MBJ_FREE, (0.1, 0.2, 0.3), 2.0, coords=[0.5, 3.0, -0.25, 0, 0, 0, 1], velocity=list(vel) + list(spin))
    mb.finish(gravity)
    return mb.scene("free_box")


def test_free_joint_link_moves_like_the_rigid_body_it_is():
    """A single FreeJoint link is a rigid body: same gyroscopic augmented mass (multibody.rs:482-490 against
    rigid_body.rs:566-586), same semi-implicit update.  20 steps of a tumbling box in free fall."""
    vel, spin = (0.3, 1.0, -0.2), (1.0, 2.5, -0.7)
    sc = _free_box(vel, spin)
    mb_out = _run(sc, 20)
    rb = abi.new_bodies(1)
    m, inertia = scenes.cuboid_mass_properties((0.1, 0.2, 0.3), 2.0)
    rb["mass"], rb["local_inertia"] = m, inertia.reshape(9)
    rb["position"][0, :3] = [0.5, 3.0, -0.25]
    rb["velocity"][0] = list(vel) + list(spin)
    o = Oracle()
    o.upload_bodies(rb)
    for k in range(20):
        o.step()
        r = o.download_body_states()[0]
        link = mb_out[k][0][1]  # body 0 is the ground, body 1 the link's record
        assert np.abs(link["position"] - r["position"]).max() < 2e-5, (k, link["position"], r["position"])
        assert np.abs(link["velocity"] - r["velocity"]).max() < 2e-4, (k, link["velocity"], r["velocity"])
    o.close()


def test_free_fall_is_semi_implicit_euler():
    sc = _free_box((0, 0, 0), (0, 0, 0))
    out = _run(sc, 10)
    dt, g = 1.0 / 60.0, -9.81
    for k in range(10):
        n = k + 1
        assert abs(out[k][1]["velocity"][0, 1] - g * dt * n) < 1e-5
        assert abs(out[k][0]["position"][1, 1] - (3.0 + g * dt * dt * n * (n + 1) / 2)) < 1e-5


def test_revolute_pendulum_small_oscillations_have_the_analytic_period():
    """One link on a revolute joint about x, centre of mass `L` from the axis, no damping: T = 2 pi sqrt(I / (m g L))
    with I the inertia about the axis."""
    L, rad, density = 0.8, 0.1, 1.0
    mb = scenes._ground_only()
    a0 = 0.05
    mb.add(-1, abi.MBJ_REVOLUTE, (rad, rad, rad), density, parent_shift=(0, 5, 0), body_shift=(0, L, 0), axis=(1, 0, 0),
           coords=[a0], damping=[0.0] * 6)
    mb.finish()
    sc = mb.scene("pendulum")
    p = abi.default_params()
    p["dt"] = 1.0 / 600.0
    out = _run(sc, 1400, params=p, f64=True)
    ang = np.array([o[1]["coords"][0, 0] for o in out])
    m, inertia = scenes.cuboid_mass_properties((rad, rad, rad), density)
    I = inertia[0, 0] + m * L * L
    T = 2 * np.pi * np.sqrt(I / (m * 9.81 * L))
    # zero crossings (downwards then upwards): half a period apart
    s = np.sign(ang)
    cross = np.nonzero(s[1:] != s[:-1])[0]
    assert len(cross) >= 2
    half = (cross[1] - cross[0]) * float(p["dt"])
    assert abs(half - T / 2) / (T / 2) < 0.01, (half, T / 2)
    assert abs(ang).max() <= a0 * 1.02  # the semi-implicit scheme does not gain energy


def test_ball_chain_without_damping_keeps_its_energy_and_its_anchor():
    """Three BallJoint links released horizontally from the world: over a second at dt = 1/600 kinetic + potential
    energy drifts by less than 1 % of the potential energy the swing converts (m g times the links' lever arms), and
    the first joint's anchor stays put (reduced coordinates cannot drift)."""
    sc = scenes.multibody_chain(abi.MBJ_BALL, links=3, damping=0.0)
    sc.mb_links["velocity"][0, :3] = [0.0, 1.0, 0.0]
    p = abi.default_params()
    p["dt"] = 1.0 / 600.0
    out = _run(sc, 600, params=p, f64=True)
    m, inertia = scenes.cuboid_mass_properties((0.2, 0.2, 0.2), 1.0)

    def energy(states):
        e = 0.0
        for b in states[1:]:
            v, w = b["velocity"][:3].astype(np.float64), b["velocity"][3:].astype(np.float64)
            e += 0.5 * m * v @ v + 0.5 * inertia[0, 0] * w @ w + m * 9.81 * float(b["position"][1])
        return e
    e0, e1 = energy(out[0][0]), energy(out[-1][0])
    assert abs(e1 - e0) < 0.01 * m * 9.81 * (0.8 + 1.6 + 2.4), (e0, e1)
    for states, links in out[::50]:
        q, t = states["position"][1, 3:], states["position"][1, :3]
        anchor = t + scenes.quat_rotate(q.astype(np.float64), np.array([0, 0, 0.8]))
        assert np.abs(anchor - [0, 5, 0]).max() < 1e-5


def test_revolute_limits_hold_and_the_motor_reaches_its_velocity():
    mb = scenes._ground_only()
    mb.add(-1, abi.MBJ_REVOLUTE, (0.1, 0.1, 0.1), 1.0, parent_shift=(0, 5, 0), body_shift=(0, 0, 0.8), axis=(1, 0, 0),
           flags=abi.MBJ_FLAG_MIN | abi.MBJ_FLAG_MAX, min_pos=-0.3, max_pos=0.2)
    mb.finish()
    mb.add(-1, abi.MBJ_REVOLUTE, (0.1, 0.1, 0.1), 1.0, parent_shift=(3, 5, 0), body_shift=(0, 0, 0.8), axis=(0, 1, 0),
           flags=abi.MBJ_FLAG_MOTOR, motor_velocity=1.5)
    mb.finish(gravity=False)
    sc = mb.scene("limits_motor")
    out = _run(sc, 240)
    ang = np.array([o[1]["coords"][0, 0] for o in out])
    assert ang.min() > -0.3 - 0.02 and ang.max() < 0.2 + 0.02, (ang.min(), ang.max())
    assert (np.abs(ang - 0.2) < 0.02).any() or (np.abs(ang + 0.3) < 0.02).any()  # it does reach a stop
    vel = np.array([o[1]["velocity"][1, 0] for o in out])
    assert abs(vel[-1] - 1.5) < 1e-3, vel[-5:]
    assert abs(out[-1][1]["impulses"][1, 0]) > 0.0  # the motor row's impulse is cached on the link


def test_prismatic_limits_stop_a_falling_slider():
    mb = scenes._ground_only()
    mb.add(-1, abi.MBJ_PRISMATIC, (0.1, 0.1, 0.1), 1.0, parent_shift=(0, 5, 0), axis=(0, 1, 0),
           flags=abi.MBJ_FLAG_MIN, min_pos=-0.5)
    mb.finish()
    out = _run(mb.scene("slider"), 120)
    off = np.array([o[1]["coords"][0, 0] for o in out])
    assert off.min() > -0.5 - 0.02
    assert abs(off[-1] + 0.5) < 0.02 and abs(out[-1][1]["velocity"][0, 0]) < 0.05


def test_a_free_box_multibody_rests_on_the_ground_like_the_rigid_box():
    """Contact rows against a multibody link (Multibody::fill_constraint_geometry, multibody.rs:971-1025): the
    FreeJoint box on the ground gets the impulses, velocities and poses of the same box as a RigidBody."""
    mb = scenes._ground_only((4.0, 0.2, 4.0))
    mb.add(-1, abi.MBJ_FREE, (0.1, 0.1, 0.1), 1.0, coords=[0.0, 0.11, 0.0, 0, 0, 0, 1])
    mb.finish()
    sc = mb.scene("box_on_ground")
    rb = scenes.boxes3(1, 1, 1)
    rb.bodies["position"][1, :3] = [0.0, 0.11, 0.0]
    gen = scenes.ContactGenerator(rb)
    o1, o2 = Oracle(), Oracle()
    o1.upload_bodies(sc.bodies)
    o1.upload_multibodies(sc.multibodies, sc.mb_links)
    o2.upload_bodies(rb.bodies)
    for k in range(30):
        m, c = gen.generate(o2.download_body_states()["position"])
        for o in (o1, o2):
            o.upload_manifolds(m, c)
            o.step()
        a, b = o1.download_body_states()[1], o2.download_body_states()[1]
        assert np.abs(a["position"] - b["position"]).max() < 1e-5, (k, a["position"], b["position"])
        assert np.abs(a["velocity"] - b["velocity"]).max() < 1e-4, (k, a["velocity"], b["velocity"])
        assert np.abs(o1.download_contact_impulses() - o2.download_contact_impulses()).max() < 1e-5
    o1.close()
    o2.close()
