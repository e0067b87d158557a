"""Measures the BASELINE.json configs other than the bench workload (1 pyramid3, 3 wall3 tall, 4 joint
chains x 10k, 5 4096 x pyramid3) in coloured mode with the CUDA-event stage timers.  Run under gpurun.

    python tools/run_configs.py [names...]     # default: all
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from nphysics_b200 import abi, scenes  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402


def tile_contacts(m, c, copies, bodies_per_copy):
    """Replicates one world's manifold set for `copies` translated worlds (keys stay unique)."""
    nm, nc = len(m), len(c)
    M = np.tile(m, copies)
    C = np.tile(c, copies)
    w = np.repeat(np.arange(copies, dtype=np.int64), nm)
    M["body1"] = (m["body1"].astype(np.int64)[None, :] + (np.arange(copies) * bodies_per_copy)[:, None]).ravel()
    M["body2"] = (m["body2"].astype(np.int64)[None, :] + (np.arange(copies) * bodies_per_copy)[:, None]).ravel()
    M["first_contact"] = (m["first_contact"].astype(np.int64)[None, :] + (np.arange(copies) * nc)[:, None]).ravel()
    wc = np.repeat(np.arange(copies, dtype=np.uint64), nc)
    C["key"] = np.tile(c["key"], copies) + wc * np.uint64(nc + 1)
    del w
    return M, C


def run(name, sc, m, c, vel, pos, steps=30, settle=40):
    p = abi.default_params()
    p["max_velocity_iterations"] = vel
    p["max_position_iterations"] = pos
    s = Solver(0)
    s.set_params(p)
    s.upload_bodies(sc.bodies)
    if len(sc.joints):
        s.upload_joints(sc.joints)
    s.upload_manifolds(m, c)
    rest = np.zeros(len(sc.bodies), dtype=abi.body_state_dtype)
    rest["position"] = sc.bodies["position"]
    rest["velocity"] = sc.bodies["velocity"]
    for _ in range(settle):
        s.step(abi.MODE_COLOURED)
        if not len(sc.joints):
            s.upload_body_states(rest)
    s.enable_timers(True)
    acc = {}
    for _ in range(steps):
        s.step(abi.MODE_COLOURED)
        for k, v in s.get_timers().items():
            acc[k] = acc.get(k, 0.0) + v
    t = {k: v / steps for k, v in acc.items()}
    st = s.get_stats()
    nb = sc.n_dynamic
    rows = int(st["n_rows_two_body"]) + int(st["n_rows_ground"])
    out = {"config": name, "bodies": nb, "manifolds": int(len(m)), "contacts": int(len(c)), "joints": int(len(sc.joints)),
           "rows": rows, "vel_iters": vel, "pos_iters": pos, "phases": int(st["n_phases_velocity"]),
           "ms_per_step": t["step"], "body_steps_per_s": nb / (t["step"] * 1e-3),
           "row_iters_per_s": rows * vel / (t["velocity_resolution"] * 1e-3) if t["velocity_resolution"] > 0 else None,
           "stage_ms": {k: round(v, 4) for k, v in t.items()}, "residual_max": float(st["residual_max"]),
           "max_penetration": float(st["max_penetration"]), "kinetic_energy": float(st["kinetic_energy"]),
           "non_finite": int(st["non_finite"])}
    print(json.dumps(out), flush=True)
    s.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["pyramid3", "wall3", "wall3_tall", "chains10k", "ragdolls10k", "pyramid3x4096"]
    for name in which:
        t0 = time.time()
        if name == "pyramid3":
            sc = scenes.pyramid3(30)
            m, c = scenes.ContactGenerator(sc).generate()
            run(name, sc, m, c, 8, 3)
        elif name == "wall3":
            sc = scenes.wall3(50, 10)
            m, c = scenes.ContactGenerator(sc).generate()
            run(name, sc, m, c, 8, 3)
        elif name == "wall3_tall":
            sc = scenes.wall3(50, 200)
            m, c = scenes.ContactGenerator(sc).generate()
            run(name, sc, m, c, 10, 5)
        elif name == "chains10k":
            # 10 000 six-link chains: 5 000 revolute + 5 000 ball, lying 2 cm above the ground so the
            # links also make contacts (mixed joint/contact rows)
            sc = scenes.joint_chains(10000, 6, kind="mixed", with_ground_collider=True, ground_y=-0.22, pitch=6.0)
            m, c = scenes.ContactGenerator(sc, search=0.0).generate()
            run(name, sc, m, c, 8, 3)
        elif name == "ragdolls10k":
            # 10 000 ragdolls (torso + head + 4 limbs, five BallConstraints each), joint rows only
            sc = scenes.ragdolls(10000)
            m = np.zeros(0, dtype=abi.manifold_dtype)
            c = np.zeros(0, dtype=abi.contact_dtype)
            run(name, sc, m, c, 8, 3)
        elif name == "pyramid3x4096":
            base = scenes.pyramid3(30)
            bm, bc = scenes.ContactGenerator(base).generate()
            copies = 4096
            sc = scenes.tile(base, copies)
            m, c = tile_contacts(bm, bc, copies, len(base.bodies))
            run(name, sc, m, c, 8, 3, steps=10, settle=16)  # 12 refinement steps of the colouring first
        print("# %s took %.1f s wall" % (name, time.time() - t0), flush=True)
