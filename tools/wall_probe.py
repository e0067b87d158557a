"""wall3 50x200 free-running: max penetration / kinetic energy every 10 steps, coloured GPU (NB2_KEMPE as set in the
environment) and, with --oracle, the sequential oracle.  The wall buckles: how noisy is the penetration metric?"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nphysics_b200 import abi, scenes  # noqa: E402
from tests.test_gpu_configs import new_oracle, new_solver, params_10_5  # noqa: E402


def run(sim, mode, steps=150):
    sc = scenes.wall3(50, 200)
    gen = scenes.ContactGenerator(sc)
    sim.set_params(params_10_5())
    sim.upload_bodies(sc.bodies)
    out = []
    pen = 0.0
    for k in range(steps):
        st = sim.download_body_states()
        m, c = gen.generate(st["position"])
        sim.upload_manifolds(m, c)
        sim.step(mode) if mode is not None else sim.step()
        if k >= 40:
            s = sim.get_stats()
            pen = max(pen, float(s["max_penetration"]))
            if k % 10 == 9:
                out.append((k + 1, round(1e3 * pen, 1), round(float(s["kinetic_energy"]), 0)))
                pen = 0.0
    return out


if "--oracle" in sys.argv:
    print("oracle", run(new_oracle(), None), flush=True)
else:
    s = new_solver()
    r = run(s, abi.MODE_COLOURED)
    ph = s.download_schedule()[0]
    print("gpu kempe=%s colours %d" % (os.environ.get("NB2_KEMPE", "1"), int(ph.max()) + 1), r, flush=True)
