"""Colour classes of a resting pile over its first 16 steps (fresh colouring, refinement steps with the Kempe stage,
cached steps), with a conflict check of every class:

  python tools/kempe_probe.py 10x8x10 50x40x50          # NB2_KEMPE=0 for the colouring without the stage
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nphysics_b200 import abi, scenes  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402


def run(grid):
    sc = bench.build_scene(grid, 10, 5)
    s = Solver(0)
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    s.upload_colliders(scenes.scene_colliders(sc))
    s.detect_pairs(scenes.LINEAR_PREDICTION)
    s.generate_manifolds()
    rest = np.zeros(len(sc.bodies), dtype=abi.body_state_dtype)
    rest["position"] = sc.bodies["position"]
    rest["velocity"] = sc.bodies["velocity"]
    for k in range(16):
        s.step(abi.MODE_COLOURED)
        s.upload_body_states(rest)
        if k in (0, 1, 2, 5, 15):
            ph, a, b = s.download_schedule()
            ok = ph >= 0
            cnt = np.bincount(ph[ok])
            bad = 0  # bodies that appear twice in one class
            for c in range(len(cnt)):
                sel = ok & (ph == c)
                bodies = np.concatenate([a[sel & (a >= 0)], b[sel & (b >= 0)]])
                bad += len(bodies) - len(np.unique(bodies))
            st = s.get_stats()
            print(grid, "step", k, "colours", len(cnt), cnt.tolist(), "conflicts", bad, "verdict",
                  int(st["schedule_verdict"]), "res", float(st["residual_max"]), flush=True)
    s.close()


if __name__ == "__main__":
    for g in sys.argv[1:]:
        run(g)
