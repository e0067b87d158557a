"""A small multibody scene for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): ragdolls standing on
the ground, two multibodies stacked on one another (a component), a pendulum between stops with a motor-driven
neighbour.  Run under gpurun: compute-sanitizer --tool <tool> python tools/sanitize_multibody.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from nphysics_b200 import abi, scenes
from nphysics_b200.solver import Solver

mb = scenes._ground_only((8.0, 0.2, 8.0))
# two ragdolls on their feet
members = scenes.multibody_ragdolls(1).mb_links
for r in range(2):
    sc1 = scenes.multibody_ragdolls(1, height=2.245, spin=0.0)
for r, x in enumerate((-3.0, -1.5)):
    root = mb.add(-1, abi.MBJ_FREE, (0.1, 0.6, 0.2), 0.3, coords=[x, 2.245, 0.0, 0, 0, 0, 1], velocity=[0.2, 0, 0.1, 0, 0, 0])
    for l in sc1.mb_links[1:]:
        he = sc1.half_extents[int(l["body"])]
        mb.add(root, abi.MBJ_BALL, tuple(he), 0.3, parent_shift=tuple(l["parent_shift"]), body_shift=tuple(l["body_shift"]))
    mb.finish()
# a component of two
mb.add(-1, abi.MBJ_FREE, (0.2, 0.1, 0.2), 1.0, coords=[1.0, 0.11, 0.0, 0, 0, 0, 1])
mb.finish()
mb.add(-1, abi.MBJ_FREE, (0.1, 0.1, 0.1), 2.0, coords=[1.05, 0.33, -0.03, 0, 0, 0, 1], velocity=[0.2, 0, 0, 0, 0, 0])
mb.finish()
# unit joints
mb.add(-1, abi.MBJ_REVOLUTE, (0.1, 0.1, 0.1), 1.0, parent_shift=(3, 3, 0), body_shift=(0, 0, 0.8), axis=(1, 0, 0),
       flags=abi.MBJ_FLAG_MIN | abi.MBJ_FLAG_MAX, min_pos=-0.3, max_pos=0.2, collider=False)
mb.finish()
w = mb.add(-1, abi.MBJ_REVOLUTE, (0.3, 0.05, 0.3), 1.0, parent_shift=(5, 3, 0), axis=(0, 1, 0), flags=abi.MBJ_FLAG_MOTOR,
           motor_velocity=1.5, motor_max_force=0.05, collider=False)
mb.add(w, abi.MBJ_FIXED, (0.05, 0.2, 0.05), 1.0, parent_shift=(0.25, 0.25, 0.0), collider=False)
mb.finish()
sc = mb.scene("sanitize_multibody")
s = Solver(0)
s.set_params(sc.params)
s.upload_bodies(sc.bodies)
s.upload_multibodies(sc.multibodies, sc.mb_links)
gen = scenes.ContactGenerator(sc)
for k in range(6):
    m, c = gen.generate(s.download_body_states()["position"])
    s.upload_manifolds(m, c)
    s.step(abi.MODE_COLOURED if k % 2 == 0 else abi.MODE_REFERENCE_ORDER)
s.synchronize()
st = s.get_stats()
print("contacts", len(c), "non_finite", int(st["non_finite"]))
s.download_multibody_links()
s.close()
print("sanitizer script done")
