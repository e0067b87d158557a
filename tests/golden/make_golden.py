"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference holds no golden vectors for this path ("parity unpinned"), and
it cannot be run here (no Rust toolchain), so these fixtures are outputs of the ORACLE on seeded
scenes.  They pin the oracle against accidental change and give the GPU tests a committed target
that does not depend on rebuilding the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from nphysics_b200 import abi, scenes  # noqa: E402


def cases():
    """name -> (scene, generator or None, params, steps)"""
    out = {}
    sc = scenes.pyramid3(6)
    out["pyramid6"] = (sc, scenes.ContactGenerator(sc), sc.params, 6)
    sc = scenes.boxes3(3, 3, 3)
    p = abi.default_params()
    p["max_velocity_iterations"] = 10
    p["max_position_iterations"] = 5
    out["boxes27_flipped"] = (sc, scenes.ContactGenerator(sc, flip_fraction=0.4), p, 5)
    sc = scenes.joint_zoo()
    out["joint_zoo"] = (sc, None, sc.params, 6)
    # revolute chains lying 2 cm above the ground: joint rows and (predictive) contact rows mixed
    sc = scenes.joint_chains(4, 6, kind="revolute", with_ground_collider=True, ground_y=-0.22)
    out["chains_on_ground"] = (sc, scenes.ContactGenerator(sc), sc.params, 5)
    return out


def run_case(solver, mode, sc, gen, params, steps, teacher=None):
    """Steps `solver`; manifolds are generated from `teacher` states when given (teacher forcing),
    else from the solver's own states.  Returns per-step states, impulses and joints."""
    solver.set_params(params)
    solver.upload_bodies(sc.bodies)
    if len(sc.joints):
        solver.upload_joints(sc.joints)
    states, impulses = [], []
    for k in range(steps):
        st = teacher["states"][k - 1] if (teacher is not None and k > 0) else solver.download_body_states()
        if teacher is not None and k > 0:
            solver.upload_body_states(st)
        if gen is not None:
            m, c = gen.generate(st["position"])
        else:
            m, c = np.zeros(0, abi.manifold_dtype), np.zeros(0, abi.contact_dtype)
        solver.upload_manifolds(m, c)
        solver.step(mode)
        solver.synchronize()
        states.append(solver.download_body_states())
        impulses.append(solver.download_contact_impulses())
    joints = solver.download_joints()
    return {"states": states, "impulses": impulses, "joints": joints}


def save(name, res):
    st = np.stack([np.concatenate([s["position"], s["velocity"]], axis=1) for s in res["states"]])
    kw = {"states": st, "joint_impulses": res["joints"]["impulses"] if len(res["joints"]) else np.zeros((0, 7), np.float32),
          "joint_broken": res["joints"]["broken"] if len(res["joints"]) else np.zeros(0, np.uint32)}
    for k, imp in enumerate(res["impulses"]):
        kw["imp%d" % k] = imp
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **kw)


def load(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    n = z["states"].shape[0]
    states = []
    for k in range(n):
        s = np.zeros(z["states"].shape[1], abi.body_state_dtype)
        s["position"] = z["states"][k][:, :7]
        s["velocity"] = z["states"][k][:, 7:]
        states.append(s)
    return {"states": states, "impulses": [z["imp%d" % k] for k in range(n)],
            "joint_impulses": z["joint_impulses"], "joint_broken": z["joint_broken"]}


if __name__ == "__main__":
    from oracle import Oracle
    for name, (sc, gen, params, steps) in cases().items():
        res = run_case(Oracle(), None, sc, gen, params, steps)
        save(name, res)
        print(name, "bodies", len(sc.bodies), "steps", steps, "contacts", [len(i) for i in res["impulses"]])
