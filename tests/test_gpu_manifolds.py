"""The device manifold producer (csrc/narrowphase.cu) and the per-step contact refresh
(nb2_update_contacts) against the host-side producer they replace (scenes.ContactGenerator, the stand-in
for ncollide's narrow phase that every other parity test feeds both the CUDA path and the oracle with).

Integer fields (bodies, counts, ids, geometry tags) must agree exactly; the device computes in f32 from
f32 poses where the host computes in f64 from the same f32 poses, so points / normals / depths agree to
f32 rounding of the world coordinates (tolerances below).  Run with -m gpu on a B200."""
import numpy as np
import pytest

from nphysics_b200 import abi, scenes
from tests.conftest import rel_err, rel_err_q

pytestmark = pytest.mark.gpu

REF, COL = abi.MODE_REFERENCE_ORDER, abi.MODE_COLOURED


def new_solver():
    from nphysics_b200.solver import Solver
    return Solver(0)


def new_oracle():
    from oracle import Oracle
    return Oracle()


def assert_same_contact_set(tag, got, want, scale):
    (mg, cg), (mw, cw) = got, want
    assert len(mg) == len(mw), (tag, len(mg), len(mw))
    assert len(cg) == len(cw), (tag, len(cg), len(cw))
    for f in ("body1", "body2", "first_contact", "num_contacts"):
        assert np.array_equal(mg[f], mw[f]), (tag, f)
    for f in ("margin1", "margin2", "friction", "restitution", "surface_velocity", "coll1_wrt_body", "coll2_wrt_body"):
        assert np.allclose(mg[f], mw[f], rtol=0, atol=1e-7), (tag, f)
    for f in ("key", "geom1", "geom2"):
        assert np.array_equal(cg[f], cw[f]), (tag, f)
    for f in ("dir1", "dir2", "dilation1", "dilation2"):
        assert np.array_equal(cg[f], cw[f]), (tag, f)
    tol = 4e-7 * scale  # a few ulps of the largest world coordinate
    for f in ("world1", "world2"):
        assert np.abs(cg[f].astype(np.float64) - cw[f]).max() <= tol, (tag, f)
    assert np.abs(cg["normal"] - cw["normal"]).max() <= 1e-6, tag
    assert np.abs(cg["depth"].astype(np.float64) - cw["depth"]).max() <= tol, tag
    for f in ("local1", "local2"):
        assert np.abs(cg[f].astype(np.float64) - cw[f]).max() <= tol, (tag, f)


def setup_device(sc, flip=0.0, search=None):
    s = new_solver()
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    if len(sc.joints):
        s.upload_joints(sc.joints)
    s.upload_colliders(scenes.scene_colliders(sc))
    s.detect_pairs(scenes.LINEAR_PREDICTION, -1.0 if search is None else search, flip)
    return s


SCENES = {
    "pyramid3": lambda: scenes.pyramid3(30),
    "boxes3_6x6x6": lambda: scenes.boxes3(6, 6, 6),
    "boxes3_20x10x20": lambda: scenes.boxes3(20, 10, 20),
    "wall3_50x10": lambda: scenes.wall3(50, 10),
    "wall3_50x200": lambda: scenes.wall3(50, 200),
    "falling_jitter": lambda: scenes.boxes3(5, 4, 5, height=0.004, jitter=0.004),
}


@pytest.mark.parametrize("name", sorted(SCENES))
@pytest.mark.parametrize("flip", [0.0, 0.5])
def test_device_pairs_and_manifolds_match_the_host_producer(name, flip):
    """Configs 1-3 (and a jittered grid whose faces overlap partially): same pairs in the same canonical
    order, same contacts, at the initial poses."""
    sc = SCENES[name]()
    gen = scenes.ContactGenerator(sc, flip_fraction=flip, order="owner")
    s = setup_device(sc, flip)
    assert s.n_pairs == gen.npairs, (name, s.n_pairs, gen.npairs)
    s.generate_manifolds()
    got = s.download_manifolds(compact=True)
    want = gen.generate()
    scale = float(np.abs(sc.bodies["position"][:, :3]).max()) + 1.0
    assert_same_contact_set(name, got[:2], want, scale)
    # the raw device layout: manifold p owns contact slots [4p, 4p + 4)
    m_raw, c_raw = s.download_manifolds()
    assert len(m_raw) == s.n_pairs and len(c_raw) == 4 * s.n_pairs
    assert np.array_equal(m_raw["first_contact"], 4 * np.arange(s.n_pairs, dtype=np.uint32))


def test_device_manifolds_follow_the_bodies():
    """Over 25 free-running coloured steps on device-produced manifolds (a jittered grid whose top layer is
    thrown upwards, part of it spinning) the device producer and the host producer, evaluated at the same
    poses with the same persistent pairs, keep giving the same contact set -- including dropped contacts."""
    sc = scenes.boxes3(6, 5, 6, height=0.008, jitter=0.001)
    top = np.nonzero(sc.bodies["position"][:, 1] > 0.9)[0]
    sc.bodies["velocity"][top, 1] = 0.6          # the top layer lifts off: its manifolds lose all contacts
    sc.bodies["velocity"][top[::3], 3] = 4.0     # and some of it tumbles: manifolds with 1-3 contacts
    gen = scenes.ContactGenerator(sc, order="owner")
    s = setup_device(sc)
    assert s.n_pairs == gen.npairs
    seen_partial = False
    for k in range(25):
        s.generate_manifolds()
        if k % 6 == 0:
            st = s.download_body_states()
            got = s.download_manifolds(compact=True)
            want = gen.generate(st["position"])
            assert_same_contact_set("step %d" % k, got[:2], want, 4.0)
            seen_partial |= bool((got[0]["num_contacts"] < 4).any()) or len(got[0]) < gen.npairs
        s.step(COL)
    assert int(s.get_stats()["non_finite"]) == 0
    assert seen_partial  # the scene did exercise manifolds with fewer than four contacts


@pytest.mark.parametrize("name,mode", [("pyramid3", REF), ("boxes3_6x6x6", REF), ("wall3_50x10", REF), ("boxes3_6x6x6", COL)])
def test_solver_on_device_manifolds_matches_oracle_on_the_same_manifolds(name, mode):
    """End to end without any contact upload: the GPU steps on the manifolds it produced itself; the oracle is
    fed the downloaded (compacted) list.  Reference order must match the oracle at 1e-5 per quantity, step by
    step (teacher-forced), exactly as with host-produced manifolds."""
    sc = SCENES[name]()
    g = setup_device(sc)
    o = new_oracle()
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    for k in range(6):
        if k:
            g.upload_body_states(o.download_body_states())
        g.generate_manifolds()
        m, c, slots = g.download_manifolds(compact=True)
        o.upload_manifolds(m, c)
        g.step(mode)
        o.step()
        g.synchronize()
        if mode == REF:
            sg, so = g.download_body_states(), o.download_body_states()
            assert rel_err_q(sg["position"], so["position"], 1e-3) <= 1e-5, k
            assert rel_err_q(sg["velocity"], so["velocity"], 1e-3) <= 1e-5, k
            ig = g.download_contact_impulses()[slots]
            assert rel_err_q(ig, o.download_contact_impulses(), 1e-6) <= 1e-5, k
    if mode == COL:
        sg, so = g.get_stats(), o.get_stats()
        assert int(sg["n_rows_two_body"]) == int(so["n_rows_two_body"])
        assert int(sg["n_rows_ground"]) == int(so["n_rows_ground"])
        assert float(sg["residual_max"]) <= 3.0 * float(so["residual_max"]) + 1e-5


def test_device_producer_is_deterministic_and_warm_starts():
    """Two contexts give identical bits (canonical pair order: the hash-grid fill order must not leak), and the
    impulse cache keeps working across steps (ids 4p + i + 1 are stable)."""
    sc = scenes.boxes3(8, 6, 8)
    outs = []
    for _ in range(2):
        s = setup_device(sc)
        for _ in range(5):
            s.generate_manifolds()
            s.step(COL)
        m, c = s.download_manifolds()
        outs.append((s.download_body_states(), s.download_contact_impulses(), m, c))
    assert np.array_equal(outs[0][0]["position"], outs[1][0]["position"])
    assert np.array_equal(outs[0][0]["velocity"], outs[1][0]["velocity"])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2].tobytes() == outs[1][2].tobytes() and outs[0][3].tobytes() == outs[1][3].tobytes()
    assert np.abs(outs[0][1]).max() > 0


def test_incremental_recolouring_keeps_the_schedule_conflict_free(monkeypatch):
    """A live scene changes its conflict graph a few groups at a time (manifolds lose and regain contacts).  The
    schedule is then edited in place -- vanished groups give their colour back, new groups are coloured by
    Jones-Plassmann rounds against the kept colour masks -- instead of being recoloured from scratch
    (stats.schedule_verdict == 3).  Every such step must still be conflict-free and cover exactly the groups
    that have rows; the simulation must stay as close to the from-scratch variant as two colourings are."""
    sc = scenes.boxes3(8, 6, 8, height=0.008, jitter=0.001)
    top = np.nonzero(sc.bodies["position"][:, 1] > 1.1)[0]
    sc.bodies["velocity"][top, 1] = 0.5
    sc.bodies["velocity"][top[::4], 3] = 3.0
    finals = []
    for incremental in ("1", "0"):
        monkeypatch.setenv("NB2_INCREMENTAL_COLOURING", incremental)  # read at nb2_create
        s = setup_device(sc)
        verdicts = []
        for k in range(40):
            s.generate_manifolds()
            s.step(COL)
            st = s.get_stats()
            verdicts.append(int(st["schedule_verdict"]))
            assert int(st["non_finite"]) == 0
            phase, a, b = s.download_schedule()
            ok = phase >= 0
            keys = np.concatenate([phase[ok & (a >= 0)].astype(np.int64) * s.n_bodies + a[ok & (a >= 0)],
                                   phase[ok & (b >= 0)].astype(np.int64) * s.n_bodies + b[ok & (b >= 0)]])
            assert len(np.unique(keys)) == len(keys), (incremental, k, verdicts[-1])
            m, _ = s.download_manifolds()
            dyn = sc.bodies["status"] == abi.BODY_DYNAMIC
            has_rows = (m["num_contacts"] > 0) & (dyn[m["body1"]] | dyn[m["body2"]])
            assert int(ok.sum()) == int(has_rows.sum()), (incremental, k)
        finals.append((s.download_body_states(), verdicts, int(st["n_phases_velocity"])))
        print("incremental=%s verdicts: %s colours at the end: %d" % (incremental, "".join(str(v) for v in verdicts), finals[-1][2]))
    assert 3 in finals[0][1] and 3 not in finals[1][1]
    assert finals[0][1].count(1) < finals[1][1].count(1)
    assert finals[0][2] <= finals[1][2] + 4        # editing in place may cost a few colours, not many
    settled = sc.bodies["position"][:, 1] < 1.0     # the layers that stay put: same pile within a tenth of a box
    assert np.abs(finals[0][0]["position"][settled, :3] - finals[1][0]["position"][settled, :3]).max() < 0.02


def test_config2_full_size_pair_and_contact_counts():
    """BASELINE config 2: 296 000 pairs / manifolds, 1 184 000 contacts, produced on the device."""
    sc = scenes.boxes3(50, 40, 50)
    s = setup_device(sc)
    assert s.n_pairs == 296000
    s.generate_manifolds()
    m, c, _ = s.download_manifolds(compact=True)
    assert (len(m), len(c)) == (296000, 1184000)
    gen = scenes.ContactGenerator(sc, order="owner")
    assert_same_contact_set("config 2", (m, c), gen.generate(), 12.0)


def test_producer_reports_misuse():
    from nphysics_b200.solver import Nb2Error
    sc = scenes.boxes3(2, 2, 2)
    s = new_solver()
    s.set_params(sc.params)
    with pytest.raises(Nb2Error) as ei:
        s.upload_colliders(scenes.scene_colliders(sc))      # bodies first
    assert ei.value.code == abi.ERR_NOT_READY
    s.upload_bodies(sc.bodies)
    with pytest.raises(Nb2Error) as ei:
        s.generate_manifolds()                              # pairs first
    assert ei.value.code == abi.ERR_NOT_READY
    bad = scenes.scene_colliders(sc)
    bad["body"][1] = 10 ** 6
    with pytest.raises(Nb2Error) as ei:
        s.upload_colliders(bad)
    assert ei.value.code == abi.ERR_BAD_INDEX
    s.upload_colliders(scenes.scene_colliders(sc))
    assert s.detect_pairs() == len(scenes.ContactGenerator(sc).a)
    s.upload_bodies(sc.bodies)                              # a new body set drops colliders and pairs
    with pytest.raises(Nb2Error):
        s.generate_manifolds()


# ------------------------------------------------------------------ nb2_update_contacts
def test_update_contacts_equals_a_full_upload():
    """Uploading only the per-step 40 bytes of every contact gives the same bits as uploading the whole
    112-byte records again (reference order and coloured)."""
    sc = scenes.boxes3(5, 5, 5)
    gen = scenes.ContactGenerator(sc)
    for mode in (REF, COL):
        full, incr = new_solver(), new_solver()
        for s in (full, incr):
            s.set_params(sc.params)
            s.upload_bodies(sc.bodies)
        m0, c0 = gen.generate()
        for k in range(6):
            st = full.download_body_states()
            m, c = gen.generate(st["position"])
            assert len(c) == len(c0)                       # a settled grid keeps its contact set
            full.upload_manifolds(m, c)
            if k == 0:
                incr.upload_manifolds(m, c)
            else:
                incr.update_contacts(abi.contact_updates_of(c))
            full.step(mode)
            incr.step(mode)
        a, b = full.download_body_states(), incr.download_body_states()
        assert np.array_equal(a["position"], b["position"]) and np.array_equal(a["velocity"], b["velocity"])
        assert np.array_equal(full.download_contact_impulses(), incr.download_contact_impulses())


def test_update_contacts_rejects_a_different_count():
    from nphysics_b200.solver import Nb2Error
    sc = scenes.boxes3(2, 2, 2)
    m, c = scenes.ContactGenerator(sc).generate()
    s = new_solver()
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    s.upload_manifolds(m, c)
    with pytest.raises(Nb2Error) as ei:
        s.update_contacts(abi.contact_updates_of(c[:-1]))
    assert ei.value.code == abi.ERR_INVALID_ARGUMENT
