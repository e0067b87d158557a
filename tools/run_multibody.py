"""Reduced-coordinate multibodies at scale: N ragdolls of examples3d/ragdoll3.rs as shipped (FreeJoint
torso + five BallJoint members, 21 dofs each), standing on the ground with their feet in contact.  GPU step time
(CUDA events) next to the oracle's single-thread time on a sample.  Run under gpurun."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nphysics_b200 import abi, scenes  # noqa: E402


def _run_gpu(sc, n, steps, contacts, stream):
    import torch
    from nphysics_b200.solver import Solver
    s = Solver(torch.cuda.current_device(), stream=stream.cuda_stream)
    s.set_params(sc.params)
    s.upload_bodies(sc.bodies)
    s.upload_multibodies(sc.multibodies, sc.mb_links)
    n_contacts = 0
    if contacts:
        s.upload_colliders(scenes.scene_colliders(sc))
        n_pairs = s.detect_pairs(scenes.LINEAR_PREDICTION)
        s.generate_manifolds()
    for _ in range(5):
        if contacts:
            s.generate_manifolds()
        s.step(abi.MODE_COLOURED)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = s.launch_count()
    e0.record()
    for _ in range(steps):
        if contacts:
            s.generate_manifolds()
        s.step(abi.MODE_COLOURED)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = s.get_stats()
    if contacts:
        m, c, _ = s.download_manifolds(compact=True)
        n_contacts = len(c)
    rec = {"ragdolls": n, "links": 6 * n, "dofs": 21 * n, "contacts": int(n_contacts), "ms_per_step": ms,
           "ragdoll_steps_per_s": n / ms * 1e3, "launches_per_step": (s.launch_count() - l0) / steps,
           "non_finite": int(st["non_finite"])}
    s.close()
    return rec


def run(n, steps, contacts, cpu_sample):
    import torch
    from nphysics_b200.solver import Solver
    sc = scenes.multibody_ragdolls(n, height=2.245 if contacts else 5.0, spin=0.0 if contacts else 2.0)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):  # the events below are recorded on the stream the library launches on
        rec = _run_gpu(sc, n, steps, contacts, stream)
    if cpu_sample:
        from oracle import Oracle
        nc = min(n, cpu_sample)
        sc2 = scenes.multibody_ragdolls(nc, height=2.245 if contacts else 5.0, spin=0.0 if contacts else 2.0)
        o = Oracle()
        o.set_params(sc2.params)
        o.upload_bodies(sc2.bodies)
        o.upload_multibodies(sc2.multibodies, sc2.mb_links)
        gen = scenes.ContactGenerator(sc2) if contacts else None
        if gen is not None:
            o.upload_manifolds(*gen.generate(o.download_body_states()["position"]))
        for _ in range(2):
            o.step()
        t0 = time.perf_counter()
        for _ in range(10):
            o.step()
        dt = (time.perf_counter() - t0) / 10
        rec["cpu_oracle"] = {"ragdolls": nc, "ms_per_step": dt * 1e3, "ragdoll_steps_per_s": nc / dt, "threads": 1}
        o.close()
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[1000, 10000])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=500)
    a = ap.parse_args()
    for n in a.n:
        for contacts in (False, True):
            print(json.dumps(run(n, a.steps, contacts, a.cpu_sample)), flush=True)
