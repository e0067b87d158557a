"""Sleeping: ActivationManager::update restated in the oracle
(src/detection/activation_manager.rs:60-201, src/object/body.rs:65-125) and its device version.

The reference has no test for this path either (parity unpinned): the CPU tests below pin the oracle
to the closed forms the algorithm implies (energy low-pass, island rule, wake-up rule); the GPU test
runs the same scenario through the C ABI and compares energies, states and impulses with the oracle.
"""
import numpy as np
import pytest

from nphysics_b200 import abi, scenes

MIX = 0.01   # mechanical_world.rs:80
THR = abi.DEFAULT_SLEEP_THRESHOLD


def new_oracle():
    from oracle import Oracle
    return Oracle()


def stacks_scene():
    """Ground + stack A (3 boxes, x=0) + stack B (3 boxes, x=2) + a lone box C (x=4): three islands."""
    rad = 0.1
    centers = []
    for x in (0.0, 2.0):
        for j in range(3):
            centers.append((x, 0.11 + 0.22 * j, 0.0))
    centers.append((4.0, 0.11, 0.0))
    bodies, he, off = scenes._make_boxes(centers, rad, 1.0, (8.0, 0.2, 8.0))
    return scenes.Scene(bodies, he, off, name="stacks")


A, B, C = [1, 2, 3], [4, 5, 6], [7]


def run_scenario(sim, mode, steps=200, wake_at=160):
    """B starts at half the default energy (sleeps first), C never sleeps, a box of B is woken at `wake_at`."""
    sc = stacks_scene()
    gen = scenes.ContactGenerator(sc)
    sim.set_params(sc.params)
    sim.upload_bodies(sc.bodies)
    act = abi.new_activation(len(sc.bodies))
    act["energy"][B] = 2.0 * THR
    act["threshold"][C] = -1.0
    sim.upload_activation(act)
    rec = []
    pos = sc.bodies["position"].copy()
    for k in range(steps):
        m, c = gen.generate(pos)
        sim.upload_manifolds(m, c)
        sim.update_activation(MIX, [B[1]] if k == wake_at else [])
        sim.step(mode)
        st = sim.download_body_states()
        a = sim.download_activation()
        stats = sim.get_stats()
        rec.append((a.copy(), st.copy(), int(stats["n_rows_two_body"]) + int(stats["n_rows_ground"]),
                    sim.download_contact_impulses().copy()))
        pos = st["position"].astype(np.float64)
    return rec


def asleep(a, idx):
    return bool(np.all(a["energy"][idx] == 0.0))


def test_energy_low_pass_and_sleep_order():
    rec = run_scenario(new_oracle(), None)
    # closed form while awake: E_k = 0.99^k E_0 + 0.01 sum 0.99^(k-j) |v_j|^2; the stacks settle by a few
    # cm/s during the first steps, which adds ~1 % (the bottom boxes barely move: within 1e-3)
    a10 = rec[9][0]
    assert np.allclose(a10["energy"][A], 4 * THR * 0.99 ** 10, rtol=2e-2)
    assert np.allclose(a10["energy"][B], 2 * THR * 0.99 ** 10, rtol=2e-2)
    assert np.isclose(a10["energy"][A[0]], 4 * THR * 0.99 ** 10, rtol=1e-3)
    assert np.all(a10["energy"][C] <= 0.04 + 1e-9) and not asleep(a10, C)
    first_sleep = {}
    for name, idx in (("A", A), ("B", B), ("C", C)):
        ks = [k for k, r in enumerate(rec) if asleep(r[0], idx)]
        first_sleep[name] = ks[0] if ks else None
    # 2*thr * 0.99^k < thr  <=>  k >= 69 ; 4*thr * 0.99^k < thr  <=>  k >= 138 (a step or two later with
    # the settling velocities: an island sleeps when its LAST body is below the threshold)
    assert 69 <= first_sleep["B"] <= 72, first_sleep
    assert 138 <= first_sleep["A"] <= 141, first_sleep
    assert first_sleep["C"] is None  # threshold None never sleeps


def test_sleeping_island_is_frozen_and_leaves_the_row_stream():
    rec = run_scenario(new_oracle(), None)
    rows_all = rec[0][2]
    rows_b_asleep = rec[100][2]
    rows_ab_asleep = rec[150][2]
    per_stack = 3 * 4 * 3  # 3 manifolds x 4 contacts x 3 rows
    assert rows_all - rows_b_asleep == per_stack
    assert rows_b_asleep - rows_ab_asleep == per_stack
    st100, st150 = rec[100][1], rec[150][1]
    assert np.all(st100["velocity"][B] == 0.0)                       # RigidBody::deactivate zeroes the velocity
    assert np.array_equal(st100["position"][B], st150["position"][B])  # and nothing integrates it any more
    assert np.any(st100["velocity"][C] != 0.0) or np.any(rec[100][3] != 0.0)  # the awake box is still solved


def test_deferred_activation_wakes_the_whole_island():
    rec = run_scenario(new_oracle(), None, steps=175, wake_at=160)
    assert asleep(rec[159][0], B)
    a = rec[160][0]
    assert not asleep(a, B)
    # the woken body got 2*thr (Body::activate, after the low-pass pass), its island mates too (:195-203)
    assert np.all(a["energy"][B] == np.float32(2 * THR))
    assert asleep(a, A)  # the neighbouring island is untouched
    # rows of the stack are back
    assert rec[160][2] - rec[159][2] == 36


def test_woken_island_is_warm_started_from_its_pre_sleep_impulses():
    """The reference's impulse cache never forgets: the pairs of a sleeping island are filtered out of the
    manifold list (mechanical_world.rs:287-300) but their cached impulses stay, so the island is warm-started
    when it wakes.  The ground contact of stack B must carry (about) the stack's weight in the very step it
    wakes -- not ramp up from zero -- and equal what it carried when it fell asleep."""
    rec = run_scenario(new_oracle(), None, steps=165, wake_at=160)
    sc = stacks_scene()
    m, c = scenes.ContactGenerator(sc).generate()
    ground_b = np.nonzero((m["body1"] == 0) & (m["body2"] == B[0]))[0]
    assert len(ground_b) == 1
    f, n = int(m["first_contact"][ground_b[0]]), int(m["num_contacts"][ground_b[0]])
    weight_dt = 3 * 0.008 * 9.81 / 60.0
    first_asleep = [k for k, r in enumerate(rec) if asleep(r[0], B)][0]
    before = rec[first_asleep - 1][3][f:f + n, 0].sum()
    during = rec[150][3][f:f + n, 0].sum()
    woken = rec[160][3][f:f + n, 0].sum()
    assert before == pytest.approx(weight_dt, rel=0.05)
    assert during == before                  # carried, step after step, while the island sleeps
    assert woken == pytest.approx(weight_dt, rel=0.05)


def test_kinematic_body_keeps_its_island_awake():
    """Kinematic bodies are island members whose energy is never updated (activation_manager.rs:81-92):
    a dynamic box in contact with one cannot sleep, an identical isolated box does."""
    o = new_oracle()
    rad = 0.1
    centers = [(0.0, 0.11, 0.0), (0.0, 0.33, 0.0), (3.0, 0.11, 0.0)]
    bodies, he, off = scenes._make_boxes(centers, rad, 1.0, (8.0, 0.2, 8.0))
    bodies["status"][2] = abi.BODY_KINEMATIC  # the upper box of the first stack
    sc = scenes.Scene(bodies, he, off, name="kin")
    gen = scenes.ContactGenerator(sc)
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    o.upload_activation(abi.new_activation(len(bodies)))
    pos = bodies["position"].copy()
    for _ in range(160):
        m, c = gen.generate(pos)
        o.upload_manifolds(m, c)
        o.update_activation(MIX)
        o.step()
        pos = o.download_body_states()["position"].astype(np.float64)
    a = o.download_activation()
    assert a["energy"][1] != 0.0 and a["energy"][1] < THR  # below the threshold, but its island holds a kinematic body
    assert a["energy"][2] == np.float32(4 * THR)           # never updated
    assert a["energy"][3] == 0.0                           # the isolated box sleeps


def test_update_requires_upload():
    o = new_oracle()
    sc = stacks_scene()
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    with pytest.raises(Exception):
        o.update_activation(MIX)


@pytest.mark.gpu
def test_gpu_sleeping_matches_oracle_reference_order():
    from nphysics_b200.solver import Solver
    g = run_scenario(Solver(0), abi.MODE_REFERENCE_ORDER, steps=175)
    o = run_scenario(new_oracle(), None, steps=175)
    for k, (rg, ro) in enumerate(zip(g, o)):
        assert np.array_equal(rg[0]["energy"], ro[0]["energy"]), "energies differ at step %d" % k
        assert np.array_equal(rg[1]["position"], ro[1]["position"]), k
        assert np.array_equal(rg[1]["velocity"], ro[1]["velocity"]), k
        assert rg[2] == ro[2], "row counts differ at step %d" % k
        assert np.array_equal(rg[3], ro[3]), k


@pytest.mark.gpu
def test_gpu_sleeping_coloured_mode():
    from nphysics_b200.solver import Solver
    rec = run_scenario(Solver(0), abi.MODE_COLOURED, steps=175)
    assert asleep(rec[100][0], B) and not asleep(rec[100][0], A)
    assert asleep(rec[150][0], A) and asleep(rec[150][0], B) and not asleep(rec[150][0], C)
    assert not asleep(rec[160][0], B)
    assert np.all(rec[150][1]["velocity"][A + B] == 0.0)
    assert rec[0][2] - rec[150][2] == 72


@pytest.mark.gpu
def test_gpu_islands_on_a_large_pile():
    """One 12x6x12 pile (a single island of 864 boxes, graph diameter ~30) in which a single body is
    above its threshold: the device label propagation must agree with the oracle's union-find."""
    from nphysics_b200.solver import Solver
    sc = scenes.boxes3(12, 6, 12)
    n0 = len(sc.bodies)
    gen = scenes.ContactGenerator(sc)
    m, c = gen.generate()
    act = abi.new_activation(n0)
    act["energy"][:] = 0.5 * THR          # everybody below the threshold ...
    act["energy"][n0 // 2] = 3.0 * THR     # ... but one box in the middle of the pile
    res = []
    for sim in (Solver(0), new_oracle()):
        sim.set_params(sc.params)
        sim.upload_bodies(sc.bodies)
        sim.upload_activation(act)
        sim.upload_manifolds(m, c)
        sim.update_activation(MIX)
        res.append(sim.download_activation())
    assert np.array_equal(res[0]["energy"], res[1]["energy"])
    assert np.all(res[1]["energy"][1:] != 0.0)  # one island, kept awake by a single body


@pytest.mark.gpu
def test_gpu_islands_and_sleep_on_a_20x20x20_pile():
    """8000 boxes in one island (graph diameter ~60): all below their thresholds -> the whole pile goes to
    sleep at once and the step that follows assembles no row; a deferred activation of one box wakes all
    8000.  Energies and states equal the oracle's."""
    from nphysics_b200.solver import Solver
    sc = scenes.boxes3(20, 20, 20)
    n = len(sc.bodies)
    gen = scenes.ContactGenerator(sc)
    m, c = gen.generate()
    act = abi.new_activation(n)
    act["energy"][:] = 0.5 * THR
    sims = (Solver(0), new_oracle())
    for sim in sims:
        sim.set_params(sc.params)
        sim.upload_bodies(sc.bodies)
        sim.upload_activation(act)
        sim.upload_manifolds(m, c)
        sim.update_activation(MIX)
        sim.step(abi.MODE_REFERENCE_ORDER)
    a_g, a_o = sims[0].download_activation(), sims[1].download_activation()
    assert np.array_equal(a_g["energy"], a_o["energy"])
    assert np.all(a_g["energy"][1:] == 0.0)
    st = sims[0].get_stats()
    assert int(st["n_rows_two_body"]) + int(st["n_rows_ground"]) == 0
    assert np.array_equal(sims[0].download_body_states()["position"], sims[1].download_body_states()["position"])
    for sim in sims:
        sim.upload_manifolds(m, c)
        sim.update_activation(MIX, [n // 2])
        sim.step(abi.MODE_REFERENCE_ORDER)
    a_g, a_o = sims[0].download_activation(), sims[1].download_activation()
    assert np.array_equal(a_g["energy"], a_o["energy"])
    assert np.all(a_g["energy"][1:] == np.float32(2 * THR))
    r2, rg = scenes.row_counts(sc, m)
    st = sims[0].get_stats()
    assert int(st["n_rows_two_body"]) == r2 and int(st["n_rows_ground"]) == rg
    sg, so = sims[0].download_body_states(), sims[1].download_body_states()
    assert np.array_equal(sg["position"], so["position"]) and np.array_equal(sg["velocity"], so["velocity"])
