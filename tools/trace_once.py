"""One settled step of the bench pile with a NB2_TRACE build: the kernels print their per-phase timelines."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nphysics_b200 import abi, scenes
from nphysics_b200.solver import Solver
sc = scenes.boxes3(50, 40, 50)
p = abi.default_params(); p["max_velocity_iterations"] = 10; p["max_position_iterations"] = 5
s = Solver(0); s.set_params(p); s.upload_bodies(sc.bodies); s.upload_colliders(scenes.scene_colliders(sc))
s.detect_pairs(scenes.LINEAR_PREDICTION); s.generate_manifolds()
rest = np.zeros(len(sc.bodies), dtype=abi.body_state_dtype)
rest["position"], rest["velocity"] = sc.bodies["position"], sc.bodies["velocity"]
for _ in range(20):
    s.step(abi.MODE_COLOURED); s.upload_body_states(rest)
s.synchronize()
