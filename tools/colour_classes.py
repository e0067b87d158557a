"""Sizes of the colour classes of the 100k pile's schedule (after the refinement passes of a cached schedule).
Built with EXTRA=-DNB2_BALANCE_ROUNDS=0 it shows the raw first-fit classes the balancing stage starts from."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nphysics_b200 import abi, scenes  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402

grid = sys.argv[1] if len(sys.argv) > 1 else "50x40x50"
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
    if k in (0, 1, 5, 15):
        ph, a, b = s.download_schedule()
        cnt = np.bincount(ph[ph >= 0])
        deg = np.bincount(np.concatenate([a[(ph >= 0) & (a >= 0)], b[(ph >= 0) & (b >= 0)]]))
        print("step", k, "colours", len(cnt), "classes", cnt.tolist(), "max body degree", int(deg.max()), flush=True)
s.close()
