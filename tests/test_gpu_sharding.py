"""Sharding on CUDA contexts: the device island labelling against the host union-find, and a
sharded run against the unsharded one.  The world_size-2 gloo tests of tests/test_sharding.py cover the host
logic with the CPU oracle; here every shard is a real nb2_context.  Run with -m gpu on a B200."""
import numpy as np
import pytest

from nphysics_b200 import abi, scenes, sharding

pytestmark = pytest.mark.gpu


def new_solver():
    from nphysics_b200.solver import Solver
    return Solver(0)


def batched_scene():
    """Five pyramids of different sizes: islands of unequal weight, each with its own ground body."""
    parts = [scenes.pyramid3(n) for n in (8, 5, 7, 3, 6)]
    off = 0
    bodies, he, co = [], [], []
    for k, p in enumerate(parts):
        b = p.bodies.copy()
        b["position"][:, 0] += 10.0 * k
        bodies.append(b)
        he.append(p.half_extents)
        co.append(p.coll_offset)
        off += len(b)
    return scenes.Scene(np.concatenate(bodies), np.concatenate(he), np.concatenate(co), name="five_pyramids")


def test_device_island_labels_equal_the_host_union_find():
    from tools.sharded_worlds import shard_scene
    sc = batched_scene()
    coll = scenes.scene_colliders(sc)
    s = new_solver()
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    s.upload_colliders(coll)
    s.detect_pairs()
    s.generate_manifolds()
    lab_dev, rows = s.label_islands()
    m, c = s.download_manifolds()
    want = sharding.island_labels(sc.bodies["status"], m["body1"], m["body2"])
    got = sharding.dense_labels(lab_dev)
    assert np.array_equal(got, want)                               # same partition, same numbering
    assert int(got.max()) + 1 == 5
    assert np.all(lab_dev[sc.bodies["status"] != abi.BODY_DYNAMIC] == -1)
    # the representative is the smallest body index of the island
    for isl in range(5):
        members = np.nonzero(got == isl)[0]
        assert np.all(lab_dev[members] == members.min())
    assert int(rows.sum()) == 3 * int(m["num_contacts"].sum())
    # bin packing by those rows: three ranks, whole islands each, loads as even as LPT makes them
    shards, lab, load = shard_scene(sc, coll, 3, 0)
    assert load.sum() == rows.sum() and load.max() <= 0.5 * load.sum()
    for sh in shards:
        dyn_ids = sh.body_ids[sc.bodies["status"][sh.body_ids] == abi.BODY_DYNAMIC]
        for isl in set(lab[dyn_ids]):                                   # islands are whole
            assert set(np.nonzero(lab == isl)[0]) <= set(sh.body_ids)


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_sharded_equals_unsharded_bit_for_bit_on_cuda(n_ranks):
    """Reference order: an island's rows keep their relative order inside a shard, and rows touch only their
    own bodies, so every shard must reproduce the unsharded bits of its bodies (poses, velocities, impulses
    summed per body) over several steps with device-produced contacts."""
    from tools.sharded_worlds import shard_scene
    sc = batched_scene()
    coll = scenes.scene_colliders(sc)
    mode = abi.MODE_REFERENCE_ORDER

    def run(bodies, colliders):
        s = new_solver()
        s.set_params(sc.params)
        s.upload_bodies(bodies)
        s.upload_colliders(colliders)
        s.detect_pairs()
        for _ in range(5):
            s.generate_manifolds()
            s.step(mode)
        return s.download_body_states(), s.get_stats()

    whole, st_whole = run(sc.bodies, coll)
    shards, lab, load = shard_scene(sc, coll, n_ranks, 0)
    rows = 0
    for sh in shards:
        part, st = run(sh.bodies, sh.localize_colliders(coll))
        assert np.array_equal(part["position"], whole["position"][sh.body_ids])
        assert np.array_equal(part["velocity"], whole["velocity"][sh.body_ids])
        rows += int(st["n_rows_two_body"]) + int(st["n_rows_ground"])
    assert rows == int(st_whole["n_rows_two_body"]) + int(st_whole["n_rows_ground"])


def test_sharded_worlds_runner_single_rank():
    """tools/sharded_worlds.py (bench.py's `sharded` record) on one rank: 12 pyramid3(6) worlds."""
    from tools.sharded_worlds import run_sharded_worlds
    rec = run_sharded_worlds(12, 0, 1, 0, None, steps=3, settle=4, base=scenes.pyramid3(6))
    assert rec["islands"] == 12 and rec["non_finite"] == 0
    assert rec["bodies"] == 12 * 21 and rec["rows"] > 0 and rec["ms_per_step"] > 0
    assert sum(rec["rows_per_rank"]) == rec["rows"]
