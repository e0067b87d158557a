"""Where does the position kernel's phase time go?  The bench scene (50x40x50 pile, rest-pose contact set) with the
position solve's tolerances varied: all contacts clean (allowed_linear_error = 1 m: barrier + cheap evaluation only)
against the default.  Run under gpurun; prints the stage timers."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nphysics_b200 import abi, scenes  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402


def run(grid, allowed, steps, pos_iters, settle):
    nx, ny, nz = [int(x) for x in grid.split("x")]
    sc = scenes.boxes3(nx, ny, nz)
    p = abi.default_params()
    p["max_velocity_iterations"] = 10
    p["max_position_iterations"] = pos_iters
    if allowed is not None:
        p["allowed_linear_error"] = allowed
    s = Solver(0)
    s.set_params(p)
    s.upload_bodies(sc.bodies)
    s.upload_colliders(scenes.scene_colliders(sc))
    s.detect_pairs(scenes.LINEAR_PREDICTION)
    s.generate_manifolds()
    rest = np.zeros(len(sc.bodies), dtype=abi.body_state_dtype)
    rest["position"], rest["velocity"] = sc.bodies["position"], sc.bodies["velocity"]
    for _ in range(settle):
        s.step(abi.MODE_COLOURED)
        s.upload_body_states(rest)
    for _ in range(5):
        s.step(abi.MODE_COLOURED)
    s.enable_timers(True)
    acc = {}
    for _ in range(steps):
        s.step(abi.MODE_COLOURED)
        for k, v in s.get_timers().items():
            acc[k] = acc.get(k, 0.0) + v / steps
    st = s.get_stats()
    s.close()
    return acc, st


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="50x40x50")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--settle", type=int, default=30)
    a = ap.parse_args()
    for allowed in (None, 1.0):
        for it in (5, 1):
            t, st = run(a.grid, allowed, a.steps, it, a.settle)
            print("allowed_linear_error", allowed, "pos_iters", it, "position_kernel %.4f velocity_kernel %.4f step %.4f phases %d"
                  % (t["position_kernel"], t["velocity_kernel"], t["step"], st["n_phases_velocity"]), flush=True)
