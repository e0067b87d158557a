"""The CPU oracle against physics invariants (the reference has no tests,
fixtures or golden vectors for this path -- "parity unpinned" -- so the restatement is pinned by
closed forms and conservation laws the reference's algorithm must satisfy)."""
import numpy as np
import pytest

from nphysics_b200 import abi, scenes
from oracle import Oracle

EMPTY_M = np.zeros(0, abi.manifold_dtype)
EMPTY_C = np.zeros(0, abi.contact_dtype)


def make(scene, params=None, f64=False):
    o = Oracle(f64=f64)
    o.set_params(params if params is not None else scene.params)
    o.upload_bodies(scene.bodies)
    if len(scene.joints):
        o.upload_joints(scene.joints)
    o.upload_manifolds(EMPTY_M, EMPTY_C)
    return o


def test_free_fall_semi_implicit_euler():
    """y_n = y0 - g dt^2 n(n+1)/2, v_n = -g n dt (moreau_jean_solver.rs:328-347, rigid_body.rs:467-505)."""
    sc = scenes.boxes3(1, 1, 1, height=3.0)
    o = make(sc)
    dt, g, n = 1.0 / 60.0, 9.81, 25
    y0 = float(sc.bodies["position"][1, 1])
    for _ in range(n):
        o.step()
    s = o.download_body_states()
    assert s["position"][1, 1] == pytest.approx(y0 - g * dt * dt * n * (n + 1) / 2.0, rel=2e-6)
    assert s["velocity"][1, 1] == pytest.approx(-g * n * dt, rel=2e-6)
    assert np.all(s["position"][0] == sc.bodies["position"][0])  # the ground never moves


def test_damping_and_velocity_caps():
    """rigid_body.rs:468-501."""
    sc = scenes.boxes3(2, 1, 1, height=3.0)
    sc.bodies["flags"][:] = 0  # no gravity
    sc.bodies["velocity"][1, :3] = (3.0, 0.0, 4.0)
    sc.bodies["velocity"][1, 3:] = (0.0, 2.0, 0.0)
    sc.bodies["linear_damping"][1] = 0.5
    sc.bodies["angular_damping"][1] = 0.25
    sc.bodies["velocity"][2, :3] = (30.0, 0.0, 40.0)
    sc.bodies["max_linear_velocity"][2] = 5.0
    sc.bodies["velocity"][2, 3:] = (0.0, 0.0, 9.0)
    sc.bodies["max_angular_velocity"][2] = 0.0
    o = make(sc)
    o.step()
    s = o.download_body_states()
    dt = 1.0 / 60.0
    assert np.allclose(s["velocity"][1, :3], np.array([3.0, 0.0, 4.0]) / (1 + dt * 0.5), rtol=1e-6)
    assert np.allclose(s["velocity"][1, 3:], np.array([0.0, 2.0, 0.0]) / (1 + dt * 0.25), rtol=1e-6)
    assert np.linalg.norm(s["velocity"][2, :3]) == pytest.approx(5.0, rel=1e-6)
    assert np.all(s["velocity"][2, 3:] == 0.0)


def test_gyroscopic_augmented_mass_matches_numpy():
    """inv_augmented_mass = (I_w + [w dt]x I_w - [I_w w dt]x)^-1 (rigid_body.rs:558-588)."""
    sc = scenes.boxes3(1, 1, 1, height=3.0)
    b = sc.bodies
    b["local_inertia"][1] = np.diag([1e-3, 2e-3, 3e-3]).reshape(9)
    ax = np.array([1.0, 2.0, 3.0]) / np.sqrt(14.0)
    ang = 0.7
    b["position"][1, 3:] = np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]])
    w = np.array([4.0, -3.0, 2.0])
    b["velocity"][1, 3:] = w
    o = make(sc)
    o.step()
    # recompute from the INITIAL state in float64
    q = b["position"][1, 3:].astype(np.float64)
    i, j, k, ww = q
    R = np.array([[ww * ww + i * i - j * j - k * k, 2 * (i * j - ww * k), 2 * (ww * j + i * k)],
                  [2 * (ww * k + i * j), ww * ww - i * i + j * j - k * k, 2 * (j * k - ww * i)],
                  [2 * (i * k - ww * j), 2 * (ww * i + j * k), ww * ww - i * i - j * j + k * k]])
    Iw = R @ np.diag([1e-3, 2e-3, 3e-3]) @ R.T
    dt = 1.0 / 60.0

    def cm(v):
        return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    aug = Iw + cm(w * dt) @ Iw - cm(Iw @ w * dt)
    inv, acc, com = o.debug_body_dynamics(1)
    assert np.allclose(inv[1:].reshape(3, 3), np.linalg.inv(aug), rtol=2e-4)
    assert inv[0] == pytest.approx(1.0 / b["mass"][1], rel=1e-6)
    assert np.allclose(acc[:3], (0, -9.81, 0), rtol=1e-6)
    assert np.allclose(acc[3:], np.linalg.inv(aug) @ (-np.cross(w, Iw @ w)), rtol=2e-4, atol=1e-6)


def test_resting_box_impulses_balance_gravity():
    """A box resting on the ground: normal impulses sum to m g dt, velocity stays ~0
    (signorini_model.rs:65-137, sor_prox.rs:181-230)."""
    sc = scenes.boxes3(1, 1, 1)
    gen = scenes.ContactGenerator(sc)
    o = make(sc)
    for _ in range(40):
        st = o.download_body_states()
        m, c = gen.generate(st["position"])
        o.upload_manifolds(m, c)
        o.step()
    imp = o.download_contact_impulses()
    assert len(imp) == 4
    assert np.all(imp[:, 0] >= 0.0)
    mass = float(sc.bodies["mass"][1])
    assert imp[:, 0].sum() == pytest.approx(mass * 9.81 / 60.0, rel=2e-3)
    s = o.download_body_states()
    assert np.abs(s["velocity"][1]).max() < 2e-3
    # the initial 0.01 overlap of the dilated shapes is pushed out to within allowed_linear_error
    assert o.get_stats()["max_penetration"] <= 0.001 + 1e-4


def test_coulomb_pyramid_bound_and_signorini_sign():
    """|lambda_t| <= mu * lambda_n per tangent and lambda_n >= 0 (sor_prox.rs:200,270-284)."""
    sc = scenes.pyramid3(8)
    sc.bodies["velocity"][1:, 0] = 0.3  # shear the pile so friction saturates somewhere
    gen = scenes.ContactGenerator(sc)
    o = make(sc)
    for _ in range(5):
        st = o.download_body_states()
        m, c = gen.generate(st["position"])
        o.upload_manifolds(m, c)
        o.step()
        imp = o.download_contact_impulses()
        assert np.all(imp[:, 0] >= 0.0)
        # friction rows of iteration k are clamped by the normal impulse of iteration k-1, and the
        # normal impulse can only be compared after its own last update: allow the last update's slack
        slack = 1e-6 + 0.5 * np.abs(imp[:, 0]).max() * 0.2
        assert np.all(np.abs(imp[:, 1]) <= 0.5 * imp[:, 0] + slack)
        assert np.all(np.abs(imp[:, 2]) <= 0.5 * imp[:, 0] + slack)
    assert (np.abs(imp[:, 1:]) > 0).any()


def test_two_body_rows_conserve_momentum():
    """Delta*WJ is equal and opposite on the two bodies of a row (sor_prox.rs:204-209)."""
    sc = scenes.boxes3(1, 2, 1, height=5.0)      # two stacked boxes far above the ground
    sc.bodies["flags"][:] = 0                      # no gravity
    sc.bodies["velocity"][1, :3] = (0.2, 1.0, 0.0)
    sc.bodies["velocity"][2, :3] = (0.0, -1.5, 0.1)
    sc.bodies["velocity"][2, 3:] = (0.3, 0.0, -0.2)
    gen = scenes.ContactGenerator(sc)
    o = make(sc)
    mass = sc.bodies["mass"][1:3].astype(np.float64)
    p0 = (mass[:, None] * sc.bodies["velocity"][1:3, :3]).sum(axis=0)
    m, c = gen.generate()
    assert len(m) == 1 and len(c) == 4
    o.upload_manifolds(m, c)
    o.step()
    s = o.download_body_states()
    p1 = (mass[:, None] * s["velocity"][1:3, :3].astype(np.float64)).sum(axis=0)
    assert np.allclose(p0, p1, atol=1e-7)
    # and the approach velocity along the normal is removed
    assert s["velocity"][2, 1] - s["velocity"][1, 1] >= -1e-4


def test_restitution_and_predictive_terms():
    """rhs += e*rhs when rhs <= -threshold; rhs += -depth*inv_dt for separated contacts
    (signorini_model.rs:92-101)."""
    sc = scenes.boxes3(1, 1, 1, height=0.005)     # 5 mm above rest: separated, predictive contact
    sc.restitution = 0.5
    sc.bodies["flags"][:] = 0
    sc.bodies["velocity"][1, 1] = -0.1             # slow: would close 1.67 mm this step, gap stays open
    gen = scenes.ContactGenerator(sc)
    o = make(sc)
    m, c = gen.generate()
    assert np.allclose(c["depth"], -0.015, atol=1e-6)   # raw depth; +0.02 margins => +0.005 overlap
    o.upload_manifolds(m, c)
    o.step()
    # dilated shapes overlap (depth' = +0.005 > 0) so no predictive term; |v| < threshold so no bounce
    s = o.download_body_states()
    assert s["velocity"][1, 1] == pytest.approx(0.0, abs=1e-4)   # 8 PGS sweeps, not a direct solve
    # fast approach: restitution kicks in, v_out = -e * v_in
    sc.bodies["velocity"][1, 1] = -2.0
    o = make(sc)
    o.upload_manifolds(m, c)
    o.step()
    s = o.download_body_states()
    assert s["velocity"][1, 1] == pytest.approx(1.0, rel=2e-3)
    # separated by more than the margins: the predictive term lets the body keep the part of its
    # velocity that closes the gap exactly
    sc2 = scenes.boxes3(1, 1, 1, height=0.0115)    # raw depth -0.0215 > -0.022 (still predicted), depth' < 0
    sc2.bodies["flags"][:] = 0
    sc2.bodies["velocity"][1, 1] = -3.0
    gen2 = scenes.ContactGenerator(sc2)
    m2, c2 = gen2.generate()
    assert len(c2) == 4 and np.all(c2["depth"] + 0.02 < 0)
    o = make(sc2)
    o.upload_manifolds(m2, c2)
    o.step()
    s = o.download_body_states()
    gap = -(float(c2["depth"][0]) + 0.02)
    assert s["velocity"][1, 1] == pytest.approx(-gap * 60.0, rel=5e-3)


def test_kinematic_body_drives_contact_rhs():
    """A kinematic body contributes J.v to the rhs only and is integrated after the solver
    (rigid_body.rs:692-699, mechanical_world.rs:328-332)."""
    sc = scenes.boxes3(1, 2, 1, height=5.0)
    sc.bodies["flags"][:] = 0
    sc.bodies["status"][1] = abi.BODY_KINEMATIC
    sc.bodies["velocity"][1, 1] = 0.5            # platform moving up
    gen = scenes.ContactGenerator(sc)
    o = make(sc)
    m, c = gen.generate()
    o.upload_manifolds(m, c)
    o.step()
    s = o.download_body_states()
    assert s["velocity"][1, 1] == pytest.approx(0.5)                 # unchanged
    assert s["position"][1, 1] == pytest.approx(sc.bodies["position"][1, 1] + 0.5 / 60.0, rel=1e-6)
    assert s["velocity"][2, 1] == pytest.approx(0.5, rel=1e-3)       # the box is carried along
    assert o.debug_row_counts().tolist() == [0, 0, 0, 8, 0, 4]       # all ground rows


def test_ball_joint_pendulum_keeps_anchor():
    sc = scenes.joint_chains(1, 3, kind="ball", with_ground_collider=False)
    o = make(sc)
    for _ in range(120):
        o.step()
    s = o.download_body_states()
    j = sc.joints
    pos = s["position"].astype(np.float64)
    for k in range(len(j)):
        b1, b2 = int(j["body1"][k]), int(j["body2"][k])
        w1 = pos[b1, :3] + scenes.quat_rotate(pos[b1, 3:], j["anchor1"][k].astype(np.float64))
        w2 = pos[b2, :3] + scenes.quat_rotate(pos[b2, 3:], j["anchor2"][k].astype(np.float64))
        assert np.linalg.norm(w1 - w2) < 0.05   # swinging chain, 8 sweeps, erp 0.2: cm-level drift
    assert o.get_stats()["non_finite"] == 0


def test_joint_zoo_stays_assembled_and_breaks_when_asked():
    sc = scenes.joint_zoo()
    o = make(sc)
    for _ in range(60):
        o.step()
    st = o.get_stats()
    assert st["non_finite"] == 0 and st["n_broken_joints"] == 0
    s = o.download_body_states()
    assert np.abs(s["position"][:, :3]).max() < 60.0
    # break thresholds (ball_constraint.rs:141-143): a tiny break force breaks every loaded joint
    sc2 = scenes.joint_zoo()
    sc2.joints["break_force_squared"] = 1e-12
    sc2.joints["break_torque_squared"] = 1e-12
    o2 = make(sc2)
    o2.step()
    assert o2.download_joints()["broken"].sum() >= len(sc2.joints) - 2
    # broken joints are skipped from the next step on (mechanical_world.rs:274-279)
    o2.step()
    assert o2.debug_row_counts()[:2].sum() <= 12


def test_pin_slot_cache_quirk_is_reproduced():
    """pin_slot_constraint.rs:222-236 stores impulse_id 2 into lin_impulses[2] and id 3 into
    ang_impulses[0]; the restatement keeps that mapping verbatim."""
    sc = scenes.joint_zoo()
    o = make(sc)
    o.step()
    j = o.download_joints()
    pin = j[j["type"] == abi.JOINT_PIN_SLOT][0]
    assert pin["impulses"][2] != 0.0 and pin["impulses"][3] != 0.0
    assert pin["impulses"][4] == 0.0


def test_scene_row_counts_match_survey():
    """SURVEY.md appendix C / section 8d."""
    sc = scenes.pyramid3(30)
    m, c = scenes.ContactGenerator(sc).generate()
    assert (sc.n_dynamic, len(m), len(c)) == (465, 1335, 5340)
    assert sum(scenes.row_counts(sc, m)) == 16020
    sc = scenes.wall3(50, 10)
    m, c = scenes.ContactGenerator(sc).generate()
    assert (sc.n_dynamic, len(m)) == (500, 50 + 450 + 490)
    sc = scenes.boxes3(5, 4, 5)
    m, c = scenes.ContactGenerator(sc).generate()
    assert len(m) == 4 * 4 * 5 + 5 * 3 * 5 + 5 * 4 * 4 + 25 and len(c) == 4 * len(m)
    mass = sc.bodies["mass"][1]
    assert mass == pytest.approx(0.008, rel=1e-6)
    assert sc.bodies["local_inertia"][1][0] == pytest.approx(5.3333e-5, rel=1e-4)


def test_f32_oracle_tracks_f64_oracle():
    sc = scenes.pyramid3(10)
    gen = scenes.ContactGenerator(sc)
    a, b = make(sc), make(sc, f64=True)
    for _ in range(10):
        st = a.download_body_states()
        m, c = gen.generate(st["position"])
        for o in (a, b):
            o.upload_body_states(st)
            o.upload_manifolds(m, c)
            o.step()
        sa, sb = a.download_body_states(), b.download_body_states()
        assert np.abs(sa["position"] - sb["position"]).max() < 1e-5
        assert np.abs(sa["velocity"] - sb["velocity"]).max() < 2e-3


def test_warm_start_cache_follows_contact_keys():
    """Impulses are carried by ContactId only (signorini_coulomb_pyramid_model.rs:104-108,233-260)."""
    sc = scenes.boxes3(2, 2, 1)
    gen = scenes.ContactGenerator(sc)
    o = make(sc)
    m, c = gen.generate()
    for _ in range(10):
        o.upload_manifolds(m, c)
        o.step()
    ref_imp = o.download_contact_impulses().copy()
    # same contacts presented in reverse manifold order with the same keys: same warm start
    order = np.arange(len(m))[::-1]
    m2 = m[order].copy()
    idx = np.concatenate([np.arange(f, f + n) for f, n in zip(m2["first_contact"], m2["num_contacts"])])
    c2 = c[idx].copy()
    m2["first_contact"] = np.concatenate([[0], np.cumsum(m2["num_contacts"])[:-1]])
    o.upload_manifolds(m2, c2)
    o.step()
    imp2 = o.download_contact_impulses()
    inv = np.empty_like(idx)
    inv[idx] = np.arange(len(idx))
    assert np.abs(imp2[inv] - ref_imp).max() < 0.05 * np.abs(ref_imp).max()
    # null keys are never cached: the first sweep starts from zero again
    c3 = c.copy()
    c3["key"] = 0
    o.clear_impulse_cache()
    o.upload_manifolds(m, c3)
    o.step()
    o.upload_manifolds(m, c3)
    o.step()
    assert o.get_stats()["residual_max"] > 0.0


def test_step_ccd_does_not_touch_the_impulse_cache():
    """step_ccd (moreau_jean_solver.rs:94-127) has no cache_impulses call: the impulses a regular step cached
    survive a CCD sub-step and warm-start the next regular step exactly as if the sub-step had not solved."""
    sc = scenes.boxes3(2, 2, 2)
    gen = scenes.ContactGenerator(sc)
    m, c = gen.generate()
    o = Oracle()
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    for _ in range(3):
        o.upload_manifolds(m, c)
        o.step()
    cached = o.download_contact_impulses().copy()
    assert np.abs(cached).max() > 0
    before = o.download_body_states().copy()
    o.upload_manifolds(m, c)
    o.step_ccd()
    assert np.array_equal(o.download_contact_impulses(), cached)
    after = o.download_body_states()
    assert not np.array_equal(after["velocity"], before["velocity"])  # the sub-step did solve and integrate
    assert np.isfinite(after["position"]).all()


def test_ragdolls_centre_of_mass_falls_freely_and_joints_hold():
    """Config 4's ragdoll topology (ragdoll3.rs:65-135 restated with BallConstraints): joint impulses are
    internal, so each figure's centre of mass follows the semi-implicit free fall y0 - g dt^2 k(k+1)/2
    while the five joints keep their anchors together."""
    sc = scenes.ragdolls(3)
    o = Oracle()
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    o.upload_joints(sc.joints)
    none_m, none_c = np.zeros(0, abi.manifold_dtype), np.zeros(0, abi.contact_dtype)
    steps = 40
    for _ in range(steps):
        o.upload_manifolds(none_m, none_c)
        o.step()
    st = o.download_body_states()
    dt, g = float(sc.params["dt"]), 9.81
    mass = sc.bodies["mass"].astype(np.float64)
    for r in range(3):
        idx = np.arange(1 + 6 * r, 7 + 6 * r)
        com0 = (mass[idx, None] * sc.bodies["position"][idx, :3]).sum(0) / mass[idx].sum()
        com = (mass[idx, None] * st["position"][idx, :3].astype(np.float64)).sum(0) / mass[idx].sum()
        assert abs(com[1] - (com0[1] - g * dt * dt * steps * (steps + 1) / 2.0)) < 2e-3
        assert abs(com[0] - com0[0]) < 2e-3 and abs(com[2] - com0[2]) < 2e-3
    J = sc.joints
    p = st["position"].astype(np.float64)
    w1 = p[J["body1"], :3] + scenes.quat_rotate(p[J["body1"], 3:7], J["anchor1"].astype(np.float64))
    w2 = p[J["body2"], :3] + scenes.quat_rotate(p[J["body2"], 3:7], J["anchor2"].astype(np.float64))
    assert np.abs(w1 - w2).max() < 5e-3
    assert int(o.get_stats()["n_rows_two_body"]) == 3 * 5 * 3


# ------------------------------------------------------------------ Line kinematics
def edge_scene():
    """Two boxes whose edges cross: A's top edge along x, B's bottom edge along z, B resting 15 mm above A
    (less than the two margins: the position solver must push them apart), plus a box whose vertex faces
    A's other top edge (Point/Line) -- contacts ncollide reports as Line/Line and Line/Point."""
    from nphysics_b200 import scenes
    rad = 0.1
    centers = [(0.0, 1.0, 0.0), (0.0, 1.0 + 2 * rad + 0.015, 2 * rad), (0.0, 1.0 + 2 * rad + 0.012, -2 * rad - 0.0)]
    bodies, he, off = scenes._make_boxes(centers, rad, 1.0, (3.0, 0.2, 3.0))
    bodies["flags"][:] = 0                      # no gravity: only the position correction moves anything
    m = np.zeros(2, dtype=abi.manifold_dtype)
    c = np.zeros(2, dtype=abi.contact_dtype)
    m["margin1"] = m["margin2"] = 0.01
    m["friction"] = 0.5
    m["coll1_wrt_body"][:, 6] = m["coll2_wrt_body"][:, 6] = 1.0
    # manifold 0: A (body 1) edge y=+rad, z=+rad along x  |  B (body 2) edge y=-rad, z=-rad along x rotated to z:
    # take B's edge x=0.. along z through (0, -rad, -rad)+... : lines cross above A's edge
    m["body1"][0], m["body2"][0], m["first_contact"][0], m["num_contacts"][0] = 1, 2, 0, 1
    c["local1"][0] = (0.03, rad, rad)
    c["dir1"][0] = (1.0, 0.0, 0.0)
    c["geom1"][0] = abi.GEOM_LINE
    c["local2"][0] = (0.0, -rad, -rad + 0.02)
    c["dir2"][0] = (0.0, 0.0, 1.0)
    c["geom2"][0] = abi.GEOM_LINE
    # manifold 1: A edge y=+rad, z=-rad along x (Line)  |  vertex of C (body 3) (Point)
    m["body1"][1], m["body2"][1], m["first_contact"][1], m["num_contacts"][1] = 1, 3, 1, 1
    c["local1"][1] = (-0.05, rad, -rad)
    c["dir1"][1] = (1.0, 0.0, 0.0)
    c["geom1"][1] = abi.GEOM_LINE
    c["local2"][1] = (rad, -rad, rad)
    c["geom2"][1] = abi.GEOM_POINT
    c["key"] = [11, 12]
    pos = bodies["position"].astype(np.float64)
    for k in range(2):                          # world points / normal / depth of the velocity rows
        b1, b2 = int(m["body1"][k]), int(m["body2"][k])
        w1 = pos[b1, :3] + c["local1"][k]
        w2 = pos[b2, :3] + c["local2"][k]
        d1 = c["dir1"][k].astype(np.float64)
        if c["geom2"][k] == abi.GEOM_LINE:
            d2 = c["dir2"][k].astype(np.float64)
            r = w1 - w2
            a, b, cc, d, e = d1 @ d1, d1 @ d2, d2 @ d2, d1 @ r, d2 @ r
            den = a * cc - b * b
            w1 = w1 + d1 * ((b * e - cc * d) / den)
            w2 = w2 + d2 * ((a * e - b * d) / den)
        else:
            w1 = w1 + d1 * (d1 @ (w2 - w1))
        n = (w2 - w1) / np.linalg.norm(w2 - w1)
        c["world1"][k], c["world2"][k], c["normal"][k] = w1, w2, n
        c["depth"][k] = -np.linalg.norm(w2 - w1)
    sc = scenes.Scene(bodies, he, off, name="edges")
    return sc, m, c


def test_line_line_and_line_point_position_correction():
    """Edge/edge and edge/vertex contacts are pushed apart until their distance is within the allowed error of
    the two margins (nonlinear_sor_prox.rs:156-309 with ContactKinematic's Line cases)."""
    sc, m, c = edge_scene()
    o = Oracle()
    p = abi.default_params()
    p["max_position_iterations"] = 30
    o.set_params(p)
    o.upload_bodies(sc.bodies)
    gap0 = np.linalg.norm(c["world2"].astype(np.float64) - c["world1"], axis=1)
    assert np.all(gap0 < 0.02 - 0.001)
    for _ in range(4):
        o.upload_manifolds(m, c)
        o.step()
    st = o.download_body_states()
    moved = np.abs(st["position"][1:4, :3] - sc.bodies["position"][1:4, :3]).max(axis=1)
    assert np.all(moved > 1e-4)                 # all three boxes were displaced
    # distance between the two features at the final poses (boxes do not rotate much: check along y)
    yA, yB, yC = st["position"][1, 1], st["position"][2, 1], st["position"][3, 1]
    assert yB - yA - 0.2 > 0.015 and yC - yA - 0.2 > 0.012
    assert float(o.get_stats()["max_penetration"]) <= 0.0011 + 1e-6


def test_signorini_model_is_frictionless_and_filters_inactive_contacts():
    """SignoriniModel as a ContactModel (signorini_model.rs:200-298): contacts with depth + margins < 0 make no
    row (is_constraint_active, :141-150), active ones make exactly one unilateral row, and a box sliding on the
    ground keeps its tangential velocity."""
    sc = scenes.boxes3(3, 1, 3)
    sc.bodies["velocity"][1:, 0] = 0.5
    gen = scenes.ContactGenerator(sc)
    m, c = gen.generate()
    c = c.copy()
    c["depth"][::2] = -0.03                     # every other contact: separated by more than the two margins
    o = Oracle()
    o.set_contact_model(1)
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    o.upload_manifolds(m, c)
    o.step()
    counts = o.debug_row_counts()               # joint bilateral (+ground), contact bilateral (+ground), unilateral (+ground)
    active = int((c["depth"] + 0.02 >= 0).sum())
    assert int(counts[4]) + int(counts[5]) == active and int(counts[2]) + int(counts[3]) == 0
    st = o.download_body_states()
    assert np.allclose(st["velocity"][1:, 0], 0.5, atol=1e-6)
    imp = o.download_contact_impulses()
    assert np.all(imp[:, 1:] == 0.0) and imp[:, 0].max() > 0.0
    o.set_contact_model(0)                      # the pyramid model brakes the same boxes
    o.upload_bodies(sc.bodies)
    o.upload_manifolds(m, c)
    o.step()
    assert np.all(o.download_body_states()["velocity"][1:, 0] < 0.5 - 1e-3)
