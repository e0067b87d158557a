"""BASELINE configs[4]: N independent pyramid3 worlds sharded over the ranks of one node.

Every rank holds the whole batched scene once, on its own GPU, only to label it: bodies + colliders go up,
the device producer finds the contact pairs, nb2_label_islands labels the islands (connected components over
dynamic bodies, csrc/activation.cu) and books every group's rows on a body.  sharding.make_shards bin-packs
whole islands onto ranks by row count (deterministic, so every rank computes the same split without talking
to the others); the rank then keeps only its own shard -- its bodies, their colliders -- in a fresh context
and steps it with contacts produced on the device.  No constraint crosses a rank, so the data path has no
collective; NCCL only gathers the per-rank nb2_stats records and the per-rank step time (max over ranks).

Used by bench.py (`sharded` record) and runnable on its own:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_worlds.py [worlds]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from nphysics_b200 import abi, scenes, sharding  # noqa: E402


def shard_scene(sc, colliders, n_ranks, device):
    """Islands and shards of a scene from the DEVICE labelling.  Returns (shards, labels, load per rank)."""
    from nphysics_b200.solver import Solver
    s = Solver(device)
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    if len(sc.joints):
        s.upload_joints(sc.joints)
    s.upload_colliders(colliders)
    s.detect_pairs(scenes.LINEAR_PREDICTION)
    s.generate_manifolds()
    labels, rows = s.label_islands()
    s.close()
    shards, lab, load = sharding.make_shards(sc.bodies, sc.joints, n_ranks=n_ranks, labels=labels, body_rows=rows)
    return shards, lab, load


def run_sharded_worlds(worlds, rank, world_size, local_rank, dist, steps=10, settle=16, base=None):
    import torch
    from nphysics_b200.solver import Solver
    base = base if base is not None else scenes.pyramid3(30)
    sc = scenes.tile(base, worlds)
    coll = scenes.scene_colliders(sc)
    shards, lab, load = shard_scene(sc, coll, world_size, local_rank)
    mine = shards[rank]
    s = Solver(local_rank)
    s.set_params(sc.params)
    s.upload_bodies(mine.bodies)
    if len(mine.joints):
        s.upload_joints(mine.joints)
    s.upload_colliders(mine.localize_colliders(coll))
    n_pairs = s.detect_pairs(scenes.LINEAR_PREDICTION)
    for _ in range(settle):  # a fresh colouring and its refinement passes
        s.generate_manifolds()
        s.step(abi.MODE_COLOURED)
    s.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    s.enable_timers(True)
    acc = 0.0
    for _ in range(steps):
        s.generate_manifolds()
        s.step(abi.MODE_COLOURED)
        acc += s.get_timers()["step"]
    ms = torch.tensor([acc / steps], device=torch.device("cuda", local_rank), dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # the job runs at the pace of its slowest rank
    stats = sharding.gather_stats(s.get_stats(), dist)
    whole = sharding.combine_stats(stats)
    s.close()
    n_dyn = int((sc.bodies["status"] == abi.BODY_DYNAMIC).sum())
    ms_v = float(ms.item())
    return {
        "config": "%d independent pyramid3 worlds, sharded by island over %d rank(s) (sharding.make_shards on the device "
                  "island labels; no data-path collective, stats over %s)" % (worlds, world_size, "NCCL" if dist is not None else "nothing"),
        "scaling": "strong", "worlds": worlds, "n_gpus": world_size, "islands": int(lab.max()) + 1,
        "bodies": n_dyn, "rows": int(whole["n_rows_two_body"]) + int(whole["n_rows_ground"]),
        "rows_per_rank": [int(x) for x in load], "pairs_this_rank": int(n_pairs),
        "ms_per_step": ms_v, "body_steps_per_s": n_dyn / (ms_v * 1e-3), "steps": steps,
        "residual_max": float(whole["residual_max"]), "max_penetration": float(whole["max_penetration"]),
        "non_finite": int(whole["non_finite"]),
    }


def main():
    import torch
    import torch.distributed as dist
    worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rec = run_sharded_worlds(worlds, rank, world_size, local_rank, dist if world_size > 1 else None)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
