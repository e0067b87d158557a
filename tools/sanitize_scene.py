import sys
sys.path.insert(0, ".")
import numpy as np
from nphysics_b200 import abi, scenes
from nphysics_b200.solver import Solver
# small mixed scene: contacts + joints, coloured and reference order, sleeping, step_ccd
sc = scenes.joint_chains(6, 4, kind="mixed", with_ground_collider=True, ground_y=-0.22, pitch=6.0)
gen = scenes.ContactGenerator(sc, search=0.0)
m, c = gen.generate()
for mode in (abi.MODE_COLOURED, abi.MODE_REFERENCE_ORDER):
    s = Solver(0)
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    s.upload_joints(sc.joints)
    s.upload_activation(abi.new_activation(len(sc.bodies)))
    for k in range(4):
        s.upload_manifolds(m, c)
        s.update_activation(0.01, [1] if k == 2 else [])
        s.step(mode)
    s.upload_manifolds(m, c)
    s.step_ccd(mode)
    st = s.get_stats()
    print(mode, int(st["n_phases_velocity"]), float(st["residual_max"]), int(st["non_finite"]))
    s.download_body_states(); s.download_contact_impulses(); s.download_joints(); s.download_activation()
    s.close()
sc = scenes.boxes3(4, 4, 4)
gen = scenes.ContactGenerator(sc)
m, c = gen.generate()
for layout in (0, 1):
    s = Solver(0)
    s.set_contact_layout(layout)
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    for k in range(16):
        s.upload_manifolds(m, c[::-1].copy() if False else c)
        s.step(abi.MODE_COLOURED)
    print("layout", layout, int(s.get_stats()["n_phases_velocity"]))
    s.close()
# device producer path: persistent pairs, manifolds re-produced every step (shared-memory staged records,
# chunk bookkeeping inside k_build_items), both orders, a free-running scene whose contacts change
sc = scenes.boxes3(5, 4, 5)
for mode in (abi.MODE_COLOURED, abi.MODE_REFERENCE_ORDER):
    s = Solver(0)
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    s.upload_colliders(scenes.scene_colliders(sc))
    print("pairs", s.detect_pairs(scenes.LINEAR_PREDICTION))
    for k in range(12):
        s.generate_manifolds()
        s.step(mode)
    st = s.get_stats()
    print("producer", mode, int(st["n_phases_velocity"]), float(st["residual_max"]), int(st["non_finite"]))
    s.download_manifolds(); s.download_body_states(); s.download_contact_impulses()
    s.close()
print("sanitizer script done")
