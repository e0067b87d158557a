"""Coloured production mode against the oracle's sequential order over a free run (manifolds regenerated from each
simulation's own poses): residual / penetration / kinetic energy averaged over the last `tail` steps, and the sink of
the highest-sinking body.  python tools/quality.py [pyramid3|wall3|boxes|pile] [steps] [tail]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from nphysics_b200 import abi, scenes  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402
from oracle import Oracle  # noqa: E402

_ORACLE_CACHE = {}


def run(name, steps, tail, with_oracle=True):
    sc = {"pyramid3": lambda: scenes.pyramid3(30), "wall3": lambda: scenes.wall3(50, 10),
          "boxes": lambda: scenes.boxes3(8, 8, 8), "pile": lambda: scenes.boxes3(12, 30, 12)}[name]()
    if name == "pile":
        sc.params["max_velocity_iterations"] = 10
        sc.params["max_position_iterations"] = 5
    gen = scenes.ContactGenerator(sc)
    sims = {"coloured": (Solver(0), abi.MODE_COLOURED)}
    if with_oracle:
        sims["oracle"] = (Oracle(), None)
    y0 = sc.bodies["position"][:, 1].astype(np.float64)
    out = {}
    for tag, (s, mode) in sims.items():
        s.set_params(sc.params)
        s.upload_bodies(sc.bodies)
        acc = {"res": [], "pen": [], "ke": []}
        for k in range(steps):
            st = s.download_body_states()
            m, c = gen.generate(st["position"])
            s.upload_manifolds(m, c)
            s.step(mode)
            if k >= steps - tail:
                stats = s.get_stats()
                acc["res"].append(float(stats["residual_max"]))
                acc["pen"].append(float(stats["max_penetration"]))
                acc["ke"].append(float(stats["kinetic_energy"]))
        y = s.download_body_states()["position"][:, 1].astype(np.float64)
        out[tag] = {"res": float(np.mean(acc["res"])), "pen_mm": 1e3 * max(acc["pen"]), "ke": float(np.mean(acc["ke"])),
                    "sink_mm": 1e3 * float((y0 - y)[1:].max()), "colours": int(stats["n_phases_velocity"])}
    return out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "pyramid3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    tail = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    r = run(name, steps, tail, with_oracle=os.environ.get("NB2_QUALITY_NO_ORACLE") is None)
    r["scene"], r["steps"] = name, steps
    r["guide"] = [os.environ.get("NB2_COLOUR_GUIDE_K", "0"), os.environ.get("NB2_COLOUR_GUIDE_STRIDE", "1")]
    print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
