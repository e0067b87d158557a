"""Developer diagnostic: convergence quality of the coloured mode vs the oracle (free running)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from nphysics_b200 import abi, scenes  # noqa: E402


def settle(make, sc, gen, steps, mode):
    s = make()
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    if len(sc.joints):
        s.upload_joints(sc.joints)
    hist = []
    for k in range(steps):
        st = s.download_body_states()
        m, c = gen.generate(st["position"])
        s.upload_manifolds(m, c)
        s.step(mode)
        hist.append(s.get_stats().copy())
    return s, hist


def summarize(tag, hist, tail=slice(40, 60)):
    res = np.mean([float(h["residual_max"]) for h in hist[tail]])
    pen = max(float(h["max_penetration"]) for h in hist[tail])
    ke = np.mean([float(h["kinetic_energy"]) for h in hist[tail]])
    print("%-28s residual %.3e penetration %.4f energy %.3e phases %d" % (tag, res, pen, ke,
                                                                      int(hist[-1]["n_phases_velocity"])))


if __name__ == "__main__":
    from nphysics_b200.solver import Solver
    from oracle import Oracle
    which = sys.argv[1] if len(sys.argv) > 1 else "pyramid3"
    sc = {"pyramid3": lambda: scenes.pyramid3(30), "wall3": lambda: scenes.wall3(50, 10),
          "boxes": lambda: scenes.boxes3(8, 8, 8), "boxes_tall": lambda: scenes.boxes3(6, 30, 6)}[which]()
    gen = scenes.ContactGenerator(sc)
    _, h = settle(Oracle, sc, gen, 60, None)
    summarize(which + " oracle", h)
    _, h = settle(lambda: Solver(0), sc, gen, 60, abi.MODE_COLOURED)
    summarize(which + " coloured", h)
    _, h = settle(lambda: Solver(0), sc, gen, 60, abi.MODE_REFERENCE_ORDER)
    summarize(which + " ref-order gpu", h)
