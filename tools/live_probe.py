"""Developer probe: the 100k pile as a live simulation (contacts re-produced on the device every step): colours,
groups per colour, row-count mix and stage times over time.  python tools/live_probe.py [steps]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from nphysics_b200 import abi, scenes  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    if len(sys.argv) > 2 and sys.argv[2].startswith("pyramids"):
        sc = scenes.tile(scenes.pyramid3(30), int(sys.argv[2][8:] or 4096))
        p = abi.default_params()
    else:
        sc = scenes.boxes3(50, 40, 50)
        p = abi.default_params()
        p["max_velocity_iterations"] = 10
        p["max_position_iterations"] = 5
    s = Solver(0)
    s.set_params(p)
    s.upload_bodies(sc.bodies)
    s.upload_colliders(scenes.scene_colliders(sc))
    s.detect_pairs(scenes.LINEAR_PREDICTION)
    s.enable_timers(True)
    for k in range(steps):
        s.generate_manifolds()
        s.step(abi.MODE_COLOURED)
        if k % 10 == 9 or k < 3:
            t = s.get_timers()
            st = s.get_stats()
            ph, a, b = s.download_schedule()
            ok = ph >= 0
            hist = np.bincount(ph[ok])
            m, _ = s.download_manifolds()
            nc = m["num_contacts"]
            print(json.dumps({"step": k, "verdict": int(st["schedule_verdict"]), "colours": int(st["n_phases_velocity"]),
                              "groups": int(ok.sum()), "groups_per_colour": [int(x) for x in hist],
                              "contacts": int(nc.sum()), "manifolds_by_contacts": [int((nc == i).sum()) for i in range(5)],
                              "ms": {k2: round(v, 3) for k2, v in t.items()},
                              "pen_mm": round(1e3 * float(st["max_penetration"]), 2), "ke": round(float(st["kinetic_energy"]), 1)}),
                  flush=True)


if __name__ == "__main__":
    main()
