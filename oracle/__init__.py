"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE: importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(nphysics_b200/) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

from nphysics_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))


def _host_stamp():
    """The oracle is compiled with -march=native: a library built on another host (the snapshot that
    travels to the GPU box carries the .so files) must be rebuilt, or it may hit an illegal instruction."""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        lines = [l for l in txt.splitlines() if l.startswith(("flags", "model name"))][:2]
        return hashlib.sha1("\n".join(lines).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False):
    """Compile liboracle.so / liboracle_f64.so with the recipe in oracle/Makefile (-O3 -march=native)."""
    targets = [os.path.join(_HERE, "liboracle.so"), os.path.join(_HERE, "liboracle_f64.so")]
    src = os.path.join(_HERE, "oracle.cpp")
    inc = os.path.join(_HERE, "multibody.inc")
    hdr = os.path.join(_HERE, "..", "include", "nphysics_b200.h")
    mk = os.path.join(_HERE, "Makefile")
    stamp_path = os.path.join(_HERE, "liboracle.host")
    stamp = _host_stamp()
    try:
        same_host = open(stamp_path).read().strip() == stamp
    except OSError:
        same_host = False
    newest = max(os.path.getmtime(src), os.path.getmtime(inc), os.path.getmtime(hdr), os.path.getmtime(mk))
    stale = force or not same_host or any((not os.path.exists(t)) or os.path.getmtime(t) < newest for t in targets)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
        with open(stamp_path, "w") as f:
            f.write(stamp + "\n")
    return targets


_libs = {}


def _load(f64=False):
    key = bool(f64)
    if key not in _libs:
        path = os.path.join(_HERE, "liboracle_f64.so" if f64 else "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.nbo_create.restype = ctypes.c_void_p
        for name in ["nbo_destroy", "nbo_set_params", "nbo_upload_bodies", "nbo_upload_body_states",
                     "nbo_upload_manifolds", "nbo_upload_joints", "nbo_clear_impulse_cache", "nbo_set_contact_model", "nbo_step",
                     "nbo_download_body_states", "nbo_download_contact_impulses", "nbo_download_joints",
                     "nbo_get_stats", "nbo_debug_body_dynamics", "nbo_debug_row_counts", "nbo_debug_mj_lambda",
                     "nbo_upload_multibodies", "nbo_download_multibody_links"]:
            getattr(lib, name).restype = ctypes.c_int
        _libs[key] = lib
    return _libs[key]


class Oracle:
    """Same call surface as nphysics_b200.solver.Solver, backed by the CPU restatement."""

    def __init__(self, f64=False):
        self.lib = _load(f64)
        self.h = ctypes.c_void_p(self.lib.nbo_create())
        self.n_bodies = 0
        self.n_contacts = 0
        self.n_joints = 0
        self.set_params(abi.default_params())

    def close(self):
        if self.h:
            self.lib.nbo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle call failed: %d" % rc)

    def set_params(self, p):
        p = np.ascontiguousarray(p, dtype=abi.params_dtype)
        self.params = p.copy()
        self._chk(self.lib.nbo_set_params(self.h, abi.ptr(p)))

    def upload_bodies(self, bodies):
        b = np.ascontiguousarray(bodies, dtype=abi.body_dtype)
        self.n_bodies = len(b)
        self._chk(self.lib.nbo_upload_bodies(self.h, abi.ptr(b), ctypes.c_uint32(len(b))))

    def upload_body_states(self, states, first=0):
        s = np.ascontiguousarray(states, dtype=abi.body_state_dtype)
        self._chk(self.lib.nbo_upload_body_states(self.h, abi.ptr(s), ctypes.c_uint32(first),
                                                  ctypes.c_uint32(len(s))))

    def upload_manifolds(self, manifolds, contacts):
        m = np.ascontiguousarray(manifolds, dtype=abi.manifold_dtype)
        c = np.ascontiguousarray(contacts, dtype=abi.contact_dtype)
        self.n_contacts = len(c)
        self._chk(self.lib.nbo_upload_manifolds(self.h, abi.ptr(m), ctypes.c_uint32(len(m)), abi.ptr(c),
                                                ctypes.c_uint32(len(c))))

    def upload_joints(self, joints):
        j = np.ascontiguousarray(joints, dtype=abi.joint_dtype)
        self.n_joints = len(j)
        self._chk(self.lib.nbo_upload_joints(self.h, abi.ptr(j), ctypes.c_uint32(len(j))))

    def clear_impulse_cache(self):
        self._chk(self.lib.nbo_clear_impulse_cache(self.h))

    # ---- reduced-coordinate multibodies (src/object/multibody.rs restated, oracle/multibody.inc)
    def upload_multibodies(self, multibodies, links):
        m = np.ascontiguousarray(multibodies, dtype=abi.multibody_dtype)
        l = np.ascontiguousarray(links, dtype=abi.mb_link_dtype)
        self.n_mb_links = len(l)
        self._chk(self.lib.nbo_upload_multibodies(self.h, abi.ptr(m), ctypes.c_uint32(len(m)), abi.ptr(l),
                                                  ctypes.c_uint32(len(l))))

    def download_multibody_links(self):
        out = np.zeros(self.n_mb_links, dtype=abi.mb_link_dtype)
        self._chk(self.lib.nbo_download_multibody_links(self.h, abi.ptr(out), ctypes.c_uint32(len(out))))
        return out

    def set_contact_model(self, model):
        """0 = SignoriniCoulombPyramidModel, 1 = SignoriniModel (frictionless)."""
        self._chk(self.lib.nbo_set_contact_model(self.h, ctypes.c_int(int(model))))

    # ---- sleeping (ActivationManager::update restated, activation_manager.rs:60-201)
    def upload_activation(self, activation):
        a = np.ascontiguousarray(activation, dtype=abi.activation_dtype)
        self._chk(self.lib.nbo_upload_activation(self.h, abi.ptr(a), ctypes.c_uint32(len(a))))

    def update_activation(self, mix_factor=0.01, to_activate=()):
        lst = np.ascontiguousarray(to_activate, dtype=np.int32)
        self._chk(self.lib.nbo_update_activation(self.h, ctypes.c_float(mix_factor),
                                                 abi.ptr(lst) if len(lst) else None, ctypes.c_uint32(len(lst))))

    def download_activation(self):
        out = np.zeros(self.n_bodies, dtype=abi.activation_dtype)
        self._chk(self.lib.nbo_download_activation(self.h, abi.ptr(out), ctypes.c_uint32(len(out))))
        return out

    def step(self, mode=None):
        self._chk(self.lib.nbo_step(self.h))

    def step_ccd(self, mode=None):
        self._chk(self.lib.nbo_step_ccd(self.h))

    def synchronize(self):
        pass

    def download_body_states(self, first=0, n=None):
        n = self.n_bodies - first if n is None else n
        out = np.zeros(n, dtype=abi.body_state_dtype)
        self._chk(self.lib.nbo_download_body_states(self.h, abi.ptr(out), ctypes.c_uint32(first),
                                                    ctypes.c_uint32(n)))
        return out

    def download_contact_impulses(self):
        out = np.zeros((self.n_contacts, 3), dtype=np.float32)
        if self.n_contacts:
            self._chk(self.lib.nbo_download_contact_impulses(self.h, abi.ptr(out),
                                                             ctypes.c_uint32(self.n_contacts)))
        return out

    def download_joints(self):
        out = np.zeros(self.n_joints, dtype=abi.joint_dtype)
        if self.n_joints:
            self._chk(self.lib.nbo_download_joints(self.h, abi.ptr(out), ctypes.c_uint32(self.n_joints)))
        return out

    def get_stats(self):
        out = np.zeros((), dtype=abi.stats_dtype)
        self._chk(self.lib.nbo_get_stats(self.h, abi.ptr(out)))
        return out

    # ---- debug taps (oracle only)
    def debug_body_dynamics(self, i):
        inv = np.zeros(10, np.float32)
        acc = np.zeros(6, np.float32)
        com = np.zeros(3, np.float32)
        self._chk(self.lib.nbo_debug_body_dynamics(self.h, ctypes.c_uint32(i), abi.ptr(inv), abi.ptr(acc),
                                                   abi.ptr(com)))
        return inv, acc, com

    def debug_row_counts(self):
        out = np.zeros(6, np.uint32)
        self._chk(self.lib.nbo_debug_row_counts(self.h, abi.ptr(out)))
        return out

    def debug_mj_lambda(self, i):
        out = np.zeros(6, np.float32)
        self._chk(self.lib.nbo_debug_mj_lambda(self.h, ctypes.c_uint32(i), abi.ptr(out)))
        return out
