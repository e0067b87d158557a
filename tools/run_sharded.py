"""BASELINE config 5 sharded over the ranks of one node: 4096 independent pyramid3 worlds, world w owned by
rank w % N (whole islands per rank, no data-path collective).  Every rank steps its own
nb2_context; the only collective is the all-gather of the nb2_stats records (nphysics_b200.sharding).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_sharded.py [worlds]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from nphysics_b200 import abi, scenes, sharding  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402
from tools.run_configs import tile_contacts  # noqa: E402


def main():
    worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mine = len(range(rank, worlds, world_size))          # worlds w with w % N == rank
    base = scenes.pyramid3(30)
    bm, bc = scenes.ContactGenerator(base).generate()
    sc = scenes.tile(base, mine)
    m, c = tile_contacts(bm, bc, mine, len(base.bodies))
    p = abi.default_params()
    s = Solver(local_rank)
    s.set_params(p)
    s.upload_bodies(sc.bodies)
    s.upload_manifolds(m, c)
    rest = np.zeros(len(sc.bodies), dtype=abi.body_state_dtype)
    rest["position"] = sc.bodies["position"]
    for _ in range(16):                                   # colouring + its 12 refinement steps
        s.step(abi.MODE_COLOURED)
        s.upload_body_states(rest)
    steps = 10
    s.synchronize()
    if world_size > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.enable_timers(True)
    torch.cuda.synchronize()
    acc = 0.0
    for _ in range(steps):
        s.step(abi.MODE_COLOURED)
        acc += s.get_timers()["step"]
    ms = torch.tensor([acc / steps], device="cuda")
    if world_size > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)         # the job runs at the pace of its slowest rank
    stats = sharding.gather_stats(s.get_stats(), dist if world_size > 1 else None)
    whole = sharding.combine_stats(stats)
    if rank == 0:
        nb = worlds * base.n_dynamic
        print(json.dumps({"config": "pyramid3 x %d sharded by world index" % worlds, "n_gpus": world_size,
                          "worlds_per_rank": mine, "bodies": nb, "ms_per_step": float(ms.item()),
                          "body_steps_per_s": nb / (float(ms.item()) * 1e-3),
                          "rows": int(whole["n_rows_two_body"]) + int(whole["n_rows_ground"]),
                          "residual_max": float(whole["residual_max"]), "non_finite": int(whole["non_finite"])}))
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
