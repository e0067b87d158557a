#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the MoreauJeanSolver hot path.

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

A "step" is one MoreauJeanSolver step (assembly + SOR-prox velocity iterations + integration +
nonlinear SOR-prox position iterations) of BASELINE.json configs[1]: the boxes3 scene scaled to a
50x40x50 = 100 000-box pile, 10 velocity + 5 position iterations, on its analytic rest-pose contact
set (input (A): 296 000 manifolds, 1 184 000 contacts, 3 552 000 rows -- the
same work every step, on both arms).  Metric: solver body-steps/s = dynamic bodies x steps / time.

  value   the contact set is resident in HBM (uploaded / produced once); K steps are timed.
  e2e     the reference-facing call sequence with HOST buffers, every step inside the timed region: the
          bodies' states go up from pinned host memory (the host owns the BodySet), the manifolds are
          produced on the device from them (nb2_generate_manifolds), the step runs, the new
          states come back.  `e2e_variants` times the two host-manifold sequences as well (full 112-byte
          records per contact, and the 40-byte per-step refresh of nb2_update_contacts).
  quality_vs_oracle  the timed scene, free-running from rest with a cold cache and fresh manifolds every
          step, coloured mode against the oracle's sequential order: residual, penetration, energy.
  sharded BASELINE configs[4]: 4096 pyramid3 worlds split over the N ranks by island
          (sharding.make_shards on the device's island labels), stats over NCCL -- strong scaling.
  uncached_ms_per_step  the timed steps again with a from-scratch colouring every step.
  live_simulation  the pile free-running with its contacts re-produced on the device every step (the schedule
          is edited in place), timed after 20 untimed free-running steps.
  multibody  SURVEY 8 f3: ragdolls of examples3d/ragdoll3.rs in reduced coordinates standing on the ground,
          split over the N ranks, with the oracle's rate on one host thread beside it.

With N > 1 `value` is N independent 100k piles, one per GPU (a single pile is ONE island: it does not
shard; "replicas", weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nphysics_b200 import abi, scenes  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", default="50x40x50", help="boxes3 grid nx x ny x nz (default = BASELINE configs[1])")
    ap.add_argument("--vel-iters", type=int, default=10)
    ap.add_argument("--pos-iters", type=int, default=5)
    ap.add_argument("--mode", default="coloured", choices=["coloured", "reference_order"])
    ap.add_argument("--cpu-sample-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-quality", action="store_true")
    ap.add_argument("--quality-steps", type=int, default=10)
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--no-multibody", action="store_true")
    ap.add_argument("--multibody-ragdolls", type=int, default=10000,
                    help="`multibody` record: ragdolls of examples3d/ragdoll3.rs as shipped (reduced coordinates, SURVEY 8 f3)")
    ap.add_argument("--sharded-worlds", type=int, default=4096)
    ap.add_argument("--contact-layout", type=int, default=0, help="0 = 100-byte row stream, 1 = compact 80-byte records")
    ap.add_argument("--no-schedule-cache", action="store_true")
    ap.add_argument("--settle", type=int, default=60, help="untimed impulse-cache settling steps at the rest pose")
    return ap.parse_args()


def build_scene(grid, vel_iters, pos_iters):
    nx, ny, nz = [int(x) for x in grid.lower().split("x")]
    sc = scenes.boxes3(nx, ny, nz)
    p = abi.default_params()
    p["max_velocity_iterations"] = vel_iters
    p["max_position_iterations"] = pos_iters
    sc.params = p
    return sc


def workload_config(args, n_bodies, n_manifolds, n_contacts, n_r2, n_rg, world):
    """The `config` object: identical on both arms (the driver compares them)."""
    return {
        "workload": "boxes3 scaled to %d boxes (%s pile at its rest pose, analytic contact set%s), %d velocity + %d "
                    "position iterations" % (n_bodies, args.grid, ", one independent world per GPU" if world > 1 else "",
                                              args.vel_iters, args.pos_iters),
        "bodies": int(n_bodies), "manifolds": int(n_manifolds), "contacts": int(n_contacts),
        "rows_two_body": int(n_r2), "rows_ground": int(n_rg),
        "l2": "row stream %.0f MB per sweep > 126 MB L2 (inputs larger than L2, no flush needed)"
              % ((100 * n_r2 + 84 * n_rg) / 1e6),
    }


def algorithmic_bytes(n_r2, n_rg, n_c, n_b, iv, ip):
    """SURVEY.md section 8(d) / BASELINE.md section 4."""
    vel = iv * (132 * n_r2 + 84 * n_rg)
    pos = ip * (96 * n_c)
    asm = 128 * n_r2 + 80 * n_rg + 96 * n_c + 64 * n_c
    return {"velocity_kernel": vel, "position_kernel": pos, "assembly": asm, "step": vel + pos + 200 * n_b + asm}


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).

    In-process NVML from a thread, every 10 ms: an `nvidia-smi -lms` child process initialises NVML inside
    the timed region and stalls this process's kernel launches while it does (measured: +9 % on a 20-step
    run).  Falls back to that child process only when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index = index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.thread = None
        self.sm, self.mx, self.reasons = [], [], set()
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()  # here, well before the timed region
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.handle = None

    def _sample(self):
        nv = self.nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        except Exception:
            mask = 0
        for nm, bit in self.BITS.items():
            if mask & bit:
                self.reasons.add(nm)

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def start(self):
        if self.handle is not None:
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm, mx, reasons = self.sm, self.mx, self.reasons
        if self.thread is not None:
            try:
                self._sample()  # at least one sample under load even for a very short region
            except Exception:
                pass
            self.stop_flag.set()
            self.thread.join(timeout=2)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.f.close()
            try:
                for line in open(self.path):
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) < 9:
                        continue
                    try:
                        sm.append(float(parts[1]))
                        mx.append(float(parts[2]))
                    except ValueError:
                        continue
                    for nm, val in zip(self.NAMES, parts[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(nm)
                os.unlink(self.path)
            except Exception:
                pass
        else:
            return out
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------
# CPU side: the oracle (the one place outside tests/ and smoke() that may execute oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_reference_run(sc, m, c, steps, warmup):
    """Times the CPU oracle on the same resident contact set (single thread: the reference has no
    threads).  Returns (seconds for `steps` steps, stats)."""
    from oracle import Oracle
    o = Oracle()
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    o.upload_manifolds(m, c)
    for _ in range(warmup):
        o.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step()
    dt = time.perf_counter() - t0
    st = o.get_stats()
    o.close()
    return dt, st


def _reference_worker(grid, vel_iters, pos_iters, steps, warmup, q):
    sc = build_scene(grid, vel_iters, pos_iters)
    m, c = scenes.ContactGenerator(sc).generate()
    dt, st = cpu_reference_run(sc, m, c, steps, warmup)
    q.put((dt, sc.n_dynamic, len(m), len(c), int(st["n_rows_two_body"]), int(st["n_rows_ground"])))


def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path on the host cores, on the SAME configuration as
    the CUDA arm (the full pile, the same contact set, the same iteration counts).  The reference is
    single-threaded by construction (no threads/rayon/SIMD in src/, one global island:
    src/world/mechanical_world.rs:263), so one world = one thread; with --gpus N the arm steps N
    independent worlds on N cores, mirroring the weak-scaling workload of the CUDA arm."""
    if rank != 0:
        return
    n_worlds = max(1, args.gpus)
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_reference_worker,
                         args=(args.grid, args.vel_iters, args.pos_iters, args.steps, args.warmup, q))
             for _ in range(n_worlds)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    dt = max(r[0] for r in res)
    nb, nm, nc, n_r2, n_rg = res[0][1:]
    value = n_worlds * nb * args.steps / dt
    line = {
        "impl": "reference", "metric": "solver body-steps/sec", "value": value, "unit": "body-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, nb, nm, nc, n_r2, n_rg, n_worlds),
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": n_worlds, "kind": "port",
                         "sample": "%d timed steps (after %d warm-up steps) of the full %s pile (%d boxes, %d contacts) "
                                   "through oracle/liboracle.so (-O3 -march=native), %d independent single-thread world(s); "
                                   "host has %d cores" % (args.steps, args.warmup, args.grid, nb, nc, n_worlds,
                                                          os.cpu_count() or 0)},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rows_per_step": n_r2 + n_rg,
    }
    print(json.dumps(line), flush=True)


def quality_vs_oracle(args, sc, device):
    """The timed scene as a free-running simulation: `quality_steps` steps from the rest pose with a cold
    impulse cache and manifolds regenerated from the current poses every step -- coloured mode on the GPU
    (device producer) against the oracle's sequential order (host producer, same canonical pair order)."""
    from nphysics_b200.solver import Solver
    from oracle import Oracle
    n = args.quality_steps
    g = Solver(device)
    g.set_params(sc.params)
    g.upload_bodies(sc.bodies)
    g.upload_colliders(scenes.scene_colliders(sc))
    g.detect_pairs(scenes.LINEAR_PREDICTION)
    for _ in range(n):
        g.generate_manifolds()
        g.step(abi.MODE_COLOURED)
    sg = g.get_stats()
    g.close()
    gen = scenes.ContactGenerator(sc, order="owner")
    o = Oracle()
    o.set_params(sc.params)
    o.upload_bodies(sc.bodies)
    for _ in range(n):
        m, c = gen.generate(o.download_body_states()["position"])
        o.upload_manifolds(m, c)
        o.step()
    so = o.get_stats()
    o.close()

    def rec(st):
        return {"residual_max": float(st["residual_max"]), "max_penetration": float(st["max_penetration"]),
                "kinetic_energy": float(st["kinetic_energy"])}

    a, b = rec(sg), rec(so)
    return {"protocol": "%d free-running steps from the rest pose, cold impulse cache, manifolds regenerated every step; "
                        "stats of the last step" % n,
            "coloured_gpu": a, "oracle_sequential": b,
            "ratio": {k: (a[k] / b[k] if b[k] else None) for k in a},
            "stated_tolerance": {"residual_max": 3.0, "max_penetration": "1.5x + 1 mm", "kinetic_energy": 4.0}}


# ----------------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    from nphysics_b200.solver import Solver
    mode = abi.MODE_COLOURED if args.mode == "coloured" else abi.MODE_REFERENCE_ORDER
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    sc = build_scene(args.grid, args.vel_iters, args.pos_iters)
    nb = sc.n_dynamic
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        solver = Solver(device=local_rank, stream=stream.cuda_stream)
        solver.set_params(sc.params)
        solver.set_contact_layout(args.contact_layout)
        solver.set_schedule_cache(not args.no_schedule_cache)
        solver.upload_bodies(sc.bodies)
        # the contact set (SURVEY 8d config 2 input (A)) is produced on the device, once, at the rest pose
        solver.upload_colliders(scenes.scene_colliders(sc))
        n_pairs = solver.detect_pairs(scenes.LINEAR_PREDICTION)
        solver.generate_manifolds()
        # settle the warm-start cache ("settled scene, warm impulse cache"): the pile
        # is held at its rest pose while the cached impulses converge
        rest = np.zeros(len(sc.bodies), dtype=abi.body_state_dtype)
        rest["position"] = sc.bodies["position"]
        rest["velocity"] = sc.bodies["velocity"]
        for _ in range(args.settle):
            solver.step(mode)
            solver.upload_body_states(rest)
        for _ in range(max(args.warmup, 3)):
            solver.step(mode)
        solver.synchronize()

        # ---------------- timed region: inputs resident in HBM, K steps
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = solver.launch_count()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            solver.step(mode)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = solver.launch_count() - l0
        clocks = sampler.stop()
        solver.synchronize()

        # ---------------- kernel times of the same step, CUDA events on the launching stream
        solver.enable_timers(True)
        acc = {}
        n_t = min(args.steps, 20)
        for _ in range(n_t):
            solver.step(mode)
            t = solver.get_timers()
            for k, v in t.items():
                acc[k] = acc.get(k, 0.0) + v
        solver.enable_timers(False)
        timers = {k: v / n_t for k, v in acc.items()}
        stats = solver.get_stats()
        n_r2, n_rg = int(stats["n_rows_two_body"]), int(stats["n_rows_ground"])
        m_dev, c_dev, _ = solver.download_manifolds(compact=True)
        n_m, n_c = len(m_dev), len(c_dev)

        # ---------------- the same K steps with a fresh colouring every step (the contact graph of a live
        # scene changes; the headline runs on an unchanged graph whose schedule is cached)
        uncached_ms = None
        if mode == abi.MODE_COLOURED and not args.no_schedule_cache:
            solver.set_schedule_cache(False)
            for _ in range(3):
                solver.step(mode)
            barrier()
            u0 = torch.cuda.Event(enable_timing=True)
            u1 = torch.cuda.Event(enable_timing=True)
            u0.record(stream)
            for _ in range(args.steps):
                solver.step(mode)
            u1.record(stream)
            barrier()
            uncached_ms = u0.elapsed_time(u1) / args.steps
            solver.set_schedule_cache(True)
            for _ in range(14):  # back to the refined cached schedule
                solver.step(mode)

        # ---------------- the same pile as a live simulation: contacts re-produced on the device from the current
        # poses every step, so the conflict graph changes whenever a manifold gains or loses its contacts
        live = None
        if mode == abi.MODE_COLOURED:
            for _ in range(20):  # untimed: the first free-running steps off the rest pose change an eighth of the
                solver.generate_manifolds()  # groups at once and colour from scratch; the record is the regime after
                solver.step(mode)
            barrier()
            v0 = torch.cuda.Event(enable_timing=True)
            v1 = torch.cuda.Event(enable_timing=True)
            v0.record(stream)
            for _ in range(args.steps):
                solver.generate_manifolds()
                solver.step(mode)
            v1.record(stream)
            barrier()
            live_ms = v0.elapsed_time(v1) / args.steps
            names = ["cached", "from_scratch", "refined", "edited_in_place"]
            hist = dict.fromkeys(names, 0)
            solver.enable_timers(True)
            lacc = {}
            n_l = min(args.steps, 20)
            for _ in range(n_l):
                solver.generate_manifolds()
                solver.step(mode)
                for k, v in solver.get_timers().items():
                    lacc[k] = lacc.get(k, 0.0) + v / n_l
                hist[names[int(solver.get_stats()["schedule_verdict"])]] += 1
            solver.enable_timers(False)
            live = {"ms_per_step": live_ms, "includes": "nb2_generate_manifolds + nb2_step per step, after 20 untimed free-running steps",
                    "schedule_verdicts_next_steps": hist, "stage_ms_next_steps": lacc,
                    "contacts_now": int(solver.download_manifolds()[0]["num_contacts"].sum())}
            # back to the rest pose and the timed contact set for the end-to-end loops
            solver.upload_body_states(rest)
            solver.generate_manifolds()
            for _ in range(5):
                solver.step(mode)
                solver.upload_body_states(rest)

        # ---------------- end to end through the C ABI with host buffers
        e2e = None
        variants = {}
        if not args.no_e2e:
            nbod = solver.n_bodies
            ssz = abi.body_state_dtype.itemsize
            ps_up = torch.empty(nbod * ssz, dtype=torch.uint8, pin_memory=True)
            ps_dn = torch.empty(nbod * ssz, dtype=torch.uint8, pin_memory=True)
            ps_up.numpy()[:] = rest.view(np.uint8).reshape(-1)
            lib, h = solver.lib, solver.h
            import ctypes
            up_ptr, dn_ptr = ctypes.c_void_p(ps_up.data_ptr()), ctypes.c_void_p(ps_dn.data_ptr())
            c_nb, c_0 = ctypes.c_uint32(nbod), ctypes.c_uint32(0)

            def timed(fn, reps):
                for _ in range(2):
                    fn()
                barrier()
                t0 = time.perf_counter()
                g0 = torch.cuda.Event(enable_timing=True)
                g1 = torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                for _ in range(reps):
                    fn()
                g1.record(stream)
                barrier()
                return max(g0.elapsed_time(g1), 1e3 * (time.perf_counter() - t0)) / reps

            def step_device_producer():
                # host BodySet -> device, contacts produced on the device, step, BodySet back to the host
                solver._chk(lib.nb2_upload_body_states(h, up_ptr, c_0, c_nb))
                solver._chk(lib.nb2_generate_manifolds(h))
                solver._chk(lib.nb2_step(h, mode))
                solver._chk(lib.nb2_download_body_states(h, dn_ptr, c_0, c_nb))

            e2e_ms = timed(step_device_producer, args.steps)
            e2e = {"ms": e2e_ms, "h2d": int(nbod * ssz), "d2h": int(nbod * ssz)}

            # host-side narrow phase instead (ncollide on the CPU): the contact set travels every step
            pm = torch.empty(max(m_dev.nbytes, 1), dtype=torch.uint8, pin_memory=True)
            pc = torch.empty(max(c_dev.nbytes, 1), dtype=torch.uint8, pin_memory=True)
            upd = abi.contact_updates_of(c_dev)
            pu = torch.empty(max(upd.nbytes, 1), dtype=torch.uint8, pin_memory=True)
            pm.numpy()[:m_dev.nbytes] = m_dev.view(np.uint8).reshape(-1)
            pc.numpy()[:c_dev.nbytes] = c_dev.view(np.uint8).reshape(-1)
            pu.numpy()[:upd.nbytes] = upd.view(np.uint8).reshape(-1)

            def step_full_upload():
                solver.upload_manifolds_raw(pm.data_ptr(), n_m, pc.data_ptr(), n_c)
                solver._chk(lib.nb2_step(h, mode))
                solver._chk(lib.nb2_download_body_states(h, dn_ptr, c_0, c_nb))

            def step_incremental():
                solver.update_contacts_raw(pu.data_ptr(), n_c)
                solver._chk(lib.nb2_step(h, mode))
                solver._chk(lib.nb2_download_body_states(h, dn_ptr, c_0, c_nb))

            reps = min(args.steps, 20)
            full_ms = timed(step_full_upload, reps)
            incr_ms = timed(step_incremental, reps)
            variants = {
                "host_manifolds_full_upload": {"ms_per_step": full_ms, "h2d_bytes_per_step": int(m_dev.nbytes + c_dev.nbytes),
                                               "d2h_bytes_per_step": int(nbod * ssz)},
                "host_manifolds_per_step_refresh": {"ms_per_step": incr_ms, "h2d_bytes_per_step": int(upd.nbytes),
                                                    "d2h_bytes_per_step": int(nbod * ssz)},
            }

    # max over ranks
    if dist is not None:
        t = torch.tensor([ms, e2e["ms"] if e2e else 0.0, uncached_ms or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        if e2e:
            e2e["ms"] = float(t[1])
        if uncached_ms is not None:
            uncached_ms = float(t[2])

    sharded = None
    if not args.no_sharded and args.mode == "coloured":
        solver.close()
        solver = None
        from tools.sharded_worlds import run_sharded_worlds
        sharded = run_sharded_worlds(args.sharded_worlds, rank, world, local_rank, dist, steps=min(args.steps, 10))

    # ragdoll3.rs as shipped -- FreeJoint torso + five BallJoint members per Multibody, feet on the ground
    # (contacts produced on the device every step); multibodies share nothing, so rank r takes every world-th ragdoll's
    # worth (strong scaling, no data-path collective: the step time is the max over ranks)
    multibody = None
    if not args.no_multibody and args.mode == "coloured":
        if solver is not None:
            solver.close()
            solver = None
        from tools.run_multibody import run as run_multibody
        n_local = (args.multibody_ragdolls + world - 1 - rank) // world
        multibody = run_multibody(n_local, 20, True, 250 if rank == 0 else 0)
        if dist is not None:
            t = torch.tensor([multibody["ms_per_step"]], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cnt = torch.tensor([float(n_local), float(multibody["contacts"])], device=dev, dtype=torch.float64)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            multibody["ragdolls_this_rank"] = n_local
            multibody["ragdolls"] = int(cnt[0])
            multibody["links"], multibody["dofs"], multibody["contacts"] = 6 * int(cnt[0]), 21 * int(cnt[0]), int(cnt[1])
            multibody["ms_per_step"] = float(t[0])
            multibody["ragdoll_steps_per_s"] = int(cnt[0]) / float(t[0]) * 1e3
            multibody["n_gpus"] = world
            multibody["scaling"] = "strong"

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak = 6650.0
            peak_src = "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
        ab = algorithmic_bytes(n_r2, n_rg, n_c, nb, args.vel_iters, args.pos_iters)
        vk_ms = timers.get("velocity_kernel", 0.0)
        achieved = ab["velocity_kernel"] / (vk_ms * 1e-3) / 1e9 if vk_ms > 0 else 0.0
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "velocity_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = tj.get("source")
            except Exception:
                traffic = None
        value = world * nb * args.steps / (ms * 1e-3)
        pk_ms = timers.get("position_kernel", 0.0)
        line = {
            "metric": "solver body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, nb, n_m, n_c, n_r2, n_rg, world),
            "b200": {"mode": args.mode, "contact_layout": args.contact_layout, "schedule_cache": not args.no_schedule_cache,
                     "contact_set": "produced on the device (nb2_detect_pairs + nb2_generate_manifolds), %d pairs" % n_pairs},
            "constraint_rows_per_sec": world * (n_r2 + n_rg) * args.vel_iters /
                                       (timers.get("velocity_resolution", 0.0) * 1e-3) if timers.get("velocity_resolution") else None,
            "roofline": {"bound": "hbm", "kernel": "k_velocity_solve_staged" if args.mode == "coloured" else "k_velocity_solve",
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": ab["velocity_kernel"], "kernel_ms": vk_ms,
                         "peak_source": peak_src,
                         "step_algorithmic_bytes": ab["step"],
                         "step_frac": ab["step"] / (ms / args.steps * 1e-3) / 1e9 / peak,
                         "position_kernel": {"algorithmic_bytes_per_launch": ab["position_kernel"], "kernel_ms": pk_ms,
                                             "frac": ab["position_kernel"] / (pk_ms * 1e-3) / 1e9 / peak if pk_ms > 0 else None}},
            "stage_ms": timers,
            "uncached_ms_per_step": uncached_ms,
            "live_simulation": live,
            "phases": {"velocity": int(stats["n_phases_velocity"]), "position": int(stats["n_phases_position"])},
            "residual_max": float(stats["residual_max"]), "max_penetration": float(stats["max_penetration"]),
            "kinetic_energy": float(stats["kinetic_energy"]), "settle_steps": args.settle,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e:
            line["e2e"] = {"value": world * nb / (e2e["ms"] * 1e-3), "unit": "body-steps/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "ms_per_step": e2e["ms"],
                           "sequence": "nb2_upload_body_states (pinned host) -> nb2_generate_manifolds -> nb2_step -> "
                                       "nb2_download_body_states (pinned host)"}
            line["e2e_variants"] = variants
        if sharded is not None:
            line["sharded"] = sharded
        if multibody is not None:
            line["multibody"] = multibody
        if world == 1 and not args.no_quality and args.mode == "coloured":
            line["quality_vs_oracle"] = quality_vs_oracle(args, sc, local_rank)
        if world == 1 and not args.no_cpu_baseline:
            mh, ch = scenes.ContactGenerator(sc).generate()
            dt, _ = cpu_reference_run(sc, mh, ch, args.cpu_sample_steps, 1)
            line["cpu_baseline"] = {"value": sc.n_dynamic * args.cpu_sample_steps / dt, "unit": "body-steps/s",
                                    "cores": 1, "kind": "port",
                                    "sample": "%d steps (after 1 warm-up step) of the full %s pile (%d boxes, %d contacts) "
                                              "through oracle/liboracle.so (-O3 -march=native), single thread (the reference "
                                              "is single-threaded); host has %d cores"
                                              % (args.cpu_sample_steps, args.grid, sc.n_dynamic, len(ch), os.cpu_count() or 0)}
        print(json.dumps(line), flush=True)
    if solver is not None:
        solver.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
