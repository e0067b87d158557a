"""Live-simulation probe: the BASELINE configs[1] pile stepped freely with the contacts re-produced on the
device every step (bench.py's `live_simulation` record), for a launch list under ncu:

  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/live.csv \
      python tools/live_probe.py --steps 4
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nphysics_b200 import abi, scenes  # noqa: E402


def main():
    import torch
    from nphysics_b200.solver import Solver
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="50x40x50")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--settle", type=int, default=30)
    a = ap.parse_args()
    sc = bench.build_scene(a.grid, 10, 5)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        s = Solver(device=0, stream=stream.cuda_stream)
        s.set_params(sc.params)
        s.upload_bodies(sc.bodies)
        s.upload_colliders(scenes.scene_colliders(sc))
        s.detect_pairs(scenes.LINEAR_PREDICTION)
        s.generate_manifolds()
        rest = np.zeros(len(sc.bodies), dtype=abi.body_state_dtype)
        rest["position"] = sc.bodies["position"]
        rest["velocity"] = sc.bodies["velocity"]
        for _ in range(a.settle):
            s.step(abi.MODE_COLOURED)
            s.upload_body_states(rest)
        for _ in range(20):  # free-running: manifolds gain and lose contacts from here on
            s.generate_manifolds()
            s.step(abi.MODE_COLOURED)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            s.generate_manifolds()
            s.step(abi.MODE_COLOURED)
        e1.record(stream)
        s.synchronize()
        print("live ms/step %.4f, verdict %d" % (e0.elapsed_time(e1) / a.steps, int(s.get_stats()["schedule_verdict"])))
        s.close()


if __name__ == "__main__":
    main()
