"""Developer diagnostic: CUDA path vs oracle on a few scenes, verbose.  Run under gpurun."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from nphysics_b200 import abi, scenes  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402
from oracle import Oracle  # noqa: E402


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b).max() if a.size else 0.0
    s = max(np.abs(b).max() if b.size else 0.0, 1e-30)
    return d, d / s


def compare(tag, so, oo):
    sg = so.download_body_states()
    sr = oo.download_body_states()
    dp, rp = rel_err(sg["position"], sr["position"])
    dv, rv = rel_err(sg["velocity"], sr["velocity"])
    ig = so.download_contact_impulses()
    ir = oo.download_contact_impulses()
    di, ri = rel_err(ig, ir)
    print("%s: pos abs %.3e rel %.3e | vel abs %.3e rel %.3e | imp abs %.3e rel %.3e" % (tag, dp, rp, dv, rv, di, ri))
    return max(rp, rv, ri)


def run_scene(sc, steps, mode, gen=None, params=None, teacher=True, verbose_every=1):
    so = Solver()
    oo = Oracle()
    p = params if params is not None else sc.params
    so.set_params(p)
    oo.set_params(p)
    so.upload_bodies(sc.bodies)
    oo.upload_bodies(sc.bodies)
    if len(sc.joints):
        so.upload_joints(sc.joints)
        oo.upload_joints(sc.joints)
    worst = 0.0
    for k in range(steps):
        st = oo.download_body_states()
        if gen is not None:
            m, c = gen.generate(st["position"])
        else:
            m = np.zeros(0, abi.manifold_dtype)
            c = np.zeros(0, abi.contact_dtype)
        if teacher:
            so.upload_body_states(st)
        so.upload_manifolds(m, c)
        oo.upload_manifolds(m, c)
        so.step(mode)
        oo.step()
        so.synchronize()
        if k % verbose_every == 0 or k == steps - 1:
            w = compare("  step %3d (nc=%d)" % (k, len(c)), so, oo)
            worst = max(worst, w)
    sg = so.get_stats()
    sr = oo.get_stats()
    for name in ["n_rows_two_body", "n_rows_ground", "n_phases_velocity", "n_phases_position", "residual_max",
                 "residual_rms", "max_penetration", "kinetic_energy"]:
        print("    %-20s gpu %-14s oracle %s" % (name, sg[name], sr[name]))
    if len(sc.joints):
        jg = so.download_joints()
        jr = oo.download_joints()
        print("    joint impulses", rel_err(jg["impulses"], jr["impulses"]), "broken", jg["broken"].sum(),
              jr["broken"].sum())
    print("  worst rel err %.3e  launches %d" % (worst, so.launch_count()))
    so.close()
    return worst


if __name__ == "__main__":
    REF, COL = abi.MODE_REFERENCE_ORDER, abi.MODE_COLOURED
    print("== free fall + spin")
    sc = scenes.boxes3(2, 1, 1, height=3.0)
    sc.bodies["velocity"][1, 3:] = (1.0, 2.0, 3.0)
    sc.bodies["velocity"][2, :3] = (0.5, 0.0, -0.2)
    run_scene(sc, 5, REF)
    print("== single box on ground (ref order)")
    sc = scenes.boxes3(1, 1, 1)
    run_scene(sc, 5, REF, scenes.ContactGenerator(sc))
    print("== 3x3x3 (ref order)")
    sc = scenes.boxes3(3, 3, 3)
    run_scene(sc, 5, REF, scenes.ContactGenerator(sc, flip_fraction=0.3))
    print("== pyramid3 (ref order, teacher forced)")
    sc = scenes.pyramid3(30)
    t = time.time()
    run_scene(sc, 6, REF, scenes.ContactGenerator(sc))
    print("   wall %.2fs" % (time.time() - t))
    print("== chains (ref order)")
    sc = scenes.joint_chains(6, 6, with_ground_collider=False)
    run_scene(sc, 6, REF)
    print("== pyramid3 (coloured, free running)")
    sc = scenes.pyramid3(30)
    run_scene(sc, 30, COL, scenes.ContactGenerator(sc), teacher=False, verbose_every=10)
    print("== chains (coloured)")
    sc = scenes.joint_chains(6, 6, with_ground_collider=False)
    run_scene(sc, 6, COL, teacher=False)
