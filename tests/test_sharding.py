"""Island labelling + island/world sharding, including the N>1 path on
world_size-2 gloo: each rank steps its shard, stats are all-gathered, and the sharded result is
bit-identical to the unsharded one (islands never exchange data inside a step)."""
import os
import socket

import numpy as np
import pytest

from nphysics_b200 import abi, scenes, sharding
from oracle import Oracle


def two_piles():
    """Two separate 2x2x2 piles + a joint chain: at least three islands sharing one ground."""
    a = scenes.boxes3(2, 2, 2)
    sc = scenes.tile(a, 2, pitch=5.0)
    return sc


def topology(sc, m):
    pa = np.concatenate([m["body1"], sc.joints["body1"]]).astype(np.int64)
    pb = np.concatenate([m["body2"], sc.joints["body2"]]).astype(np.int64)
    w = np.concatenate([3 * m["num_contacts"].astype(np.int64), np.full(len(sc.joints), 5, dtype=np.int64)])
    return pa, pb, w


def test_island_labels_skip_static_bodies():
    sc = two_piles()
    m, c = scenes.ContactGenerator(sc).generate()
    pa, pb, w = topology(sc, m)
    lab = sharding.island_labels(sc.bodies["status"], pa, pb)
    assert lab[0] == -1 and lab[9] == -1                 # the two ground bodies
    assert len(set(lab[lab >= 0])) == 2                  # two piles, not glued by the ground
    assert len(set(lab[1:9])) == 1 and len(set(lab[10:18])) == 1
    assert lab[1] != lab[10]


def test_bin_packing_balances_rows():
    rank_of, load = sharding.assign_islands([100, 90, 50, 40, 10, 10], 2)
    assert load.sum() == 300 and abs(int(load[0]) - int(load[1])) <= 20
    rank_of2, _ = sharding.assign_islands([100, 90, 50, 40, 10, 10], 2)
    assert np.all(rank_of == rank_of2)                   # deterministic


def run_unsharded(sc, m, c, steps):
    o = Oracle()
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    for _ in range(steps):
        o.upload_manifolds(m, c)
        o.step()
    return o.download_body_states(), o.get_stats()


def run_shard(shard, sc, m, c, steps):
    o = Oracle()
    o.set_params(sc.params)
    o.upload_bodies(shard.bodies)
    if len(shard.joints):
        o.upload_joints(shard.joints)
    lm, lc, _ = shard.localize_manifolds(m, c)
    for _ in range(steps):
        o.upload_manifolds(lm, lc)
        o.step()
    return o.download_body_states(), o.get_stats()


def test_sharded_equals_unsharded_bitwise():
    sc = two_piles()
    m, c = scenes.ContactGenerator(sc).generate()
    pa, pb, w = topology(sc, m)
    shards, lab, load = sharding.make_shards(sc.bodies, sc.joints, pa, pb, w, 2)
    assert load.min() > 0
    full, full_stats = run_unsharded(sc, m, c, 4)
    rows = 0
    for sh in shards:
        st, stats = run_shard(sh, sc, m, c, 4)
        dyn = sc.bodies["status"][sh.body_ids] == abi.BODY_DYNAMIC
        assert np.array_equal(st["position"][dyn], full["position"][sh.body_ids][dyn])
        assert np.array_equal(st["velocity"][dyn], full["velocity"][sh.body_ids][dyn])
        rows += int(stats["n_rows_two_body"]) + int(stats["n_rows_ground"])
    assert rows == int(full_stats["n_rows_two_body"]) + int(full_stats["n_rows_ground"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = two_piles()
        m, c = scenes.ContactGenerator(sc).generate()
        pa, pb, w = topology(sc, m)
        shards, _, _ = sharding.make_shards(sc.bodies, sc.joints, pa, pb, w, world)
        st, stats = run_shard(shards[rank], sc, m, c, 3)
        allst = sharding.gather_stats(stats, dist)
        total = sharding.combine_stats(allst)
        q.put((rank, int(total["n_rows_two_body"]) + int(total["n_rows_ground"]), float(total["residual_max"]),
               int(stats["n_rows_two_body"]) + int(stats["n_rows_ground"]), len(allst)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_world_gathers_stats():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = two_piles()
    m, c = scenes.ContactGenerator(sc).generate()
    _, full_stats = run_unsharded(sc, m, c, 3)
    full_rows = int(full_stats["n_rows_two_body"]) + int(full_stats["n_rows_ground"])
    res.sort()
    assert res[0][1] == res[1][1] == full_rows              # every rank sees the whole-job total
    assert res[0][3] + res[1][3] == full_rows               # and owns a disjoint part of it
    assert res[0][4] == 2
    assert max(res[0][2], res[1][2]) == pytest.approx(float(full_stats["residual_max"]), rel=1e-6)
