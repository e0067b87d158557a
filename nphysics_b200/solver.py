"""ctypes binding of libnphysics_b200.so (the C ABI in include/nphysics_b200.h).

This is the harness-side view of the boundary: tests and bench.py drive the CUDA path through
exactly the entry points a Rust/C++ host would bind.  There is no CPU fallback here: if the
library is missing or no sm_100 device is present the constructor raises.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnphysics_b200.so")

EXPORTS = [
    "nb2_abi_version", "nb2_error_string", "nb2_default_params", "nb2_sizeof", "nb2_combine_materials",
    "nb2_create", "nb2_destroy", "nb2_last_error", "nb2_set_params", "nb2_get_params", "nb2_enable_timers",
    "nb2_set_schedule_cache", "nb2_set_contact_layout",
    "nb2_upload_bodies", "nb2_upload_body_states", "nb2_upload_manifolds", "nb2_upload_joints",
    "nb2_clear_impulse_cache", "nb2_upload_activation", "nb2_update_activation", "nb2_download_activation",
    "nb2_step", "nb2_step_ccd", "nb2_synchronize", "nb2_download_body_states",
    "nb2_download_contact_impulses", "nb2_download_joints", "nb2_get_stats", "nb2_get_timers",
    "nb2_launch_count", "nb2_download_schedule",
    "nb2_update_contacts", "nb2_upload_colliders", "nb2_detect_pairs", "nb2_generate_manifolds",
    "nb2_download_manifolds", "nb2_label_islands", "nb2_set_contact_model",
    "nb2_upload_multibodies", "nb2_download_multibody_links",
]


def build(force=False, extra=""):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "nphysics_b200.h"))
    stale = force or not os.path.exists(LIB_PATH) or \
        os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs)
    if stale:
        cmd = ["make", "-C", src_dir, "-s", "-j4"]
        if force:
            cmd.append("-B")
        if extra:
            cmd.append("EXTRA=" + extra)
        subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load():
    """dlopen the library and declare signatures; raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libnphysics_b200.so is not built (run __graft_entry__.build()); "
                           "the CUDA path has no fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.nb2_error_string.restype = ctypes.c_char_p
    lib.nb2_last_error.restype = ctypes.c_char_p
    lib.nb2_last_error.argtypes = [ctypes.c_void_p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("nb2_error_string", "nb2_last_error"):
            fn.restype = ctypes.c_int
    lib.nb2_create.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    for i, d in enumerate(abi.SIZEOF_ORDER):
        got = lib.nb2_sizeof(i)
        if got != d.itemsize:
            raise RuntimeError("ABI struct %d size mismatch: library %d, binding %d" % (i, got, d.itemsize))
    _lib = lib
    return lib


class Nb2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nb2 error %d: %s" % (code, msg))
        self.code = code


class Solver:
    """One nb2_context: a world's solver state on one GPU."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = ctypes.c_void_p()
        rc = self.lib.nb2_create(int(device), ctypes.c_void_p(stream) if stream else None, ctypes.byref(h))
        if rc != 0:
            msg = self.lib.nb2_last_error(None)
            raise Nb2Error(rc, msg.decode() if msg else self.lib.nb2_error_string(rc).decode())
        self.h = h
        self.n_bodies = 0
        self.n_contacts = 0
        self.n_joints = 0
        self._keep = []  # host arrays of asynchronous uploads

    def close(self):
        if getattr(self, "h", None):
            self.lib.nb2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            msg = self.lib.nb2_last_error(self.h)
            raise Nb2Error(rc, msg.decode() if msg else "?")

    def set_params(self, p):
        p = np.ascontiguousarray(p, dtype=abi.params_dtype)
        self._chk(self.lib.nb2_set_params(self.h, abi.ptr(p)))

    def get_params(self):
        p = np.zeros((), dtype=abi.params_dtype)
        self._chk(self.lib.nb2_get_params(self.h, abi.ptr(p)))
        return p

    def enable_timers(self, on=True):
        self._chk(self.lib.nb2_enable_timers(self.h, 1 if on else 0))

    def set_schedule_cache(self, on=True):
        self._chk(self.lib.nb2_set_schedule_cache(self.h, 1 if on else 0))

    def set_contact_layout(self, layout):
        self._chk(self.lib.nb2_set_contact_layout(self.h, int(layout)))

    def upload_bodies(self, bodies):
        b = np.ascontiguousarray(bodies, dtype=abi.body_dtype)
        self.n_bodies = len(b)
        self._chk(self.lib.nb2_upload_bodies(self.h, abi.ptr(b), ctypes.c_uint32(len(b))))

    def upload_body_states(self, states, first=0):
        s = np.ascontiguousarray(states, dtype=abi.body_state_dtype)
        self._chk(self.lib.nb2_upload_body_states(self.h, abi.ptr(s), ctypes.c_uint32(first),
                                                  ctypes.c_uint32(len(s))))

    def upload_manifolds(self, manifolds, contacts):
        m = np.ascontiguousarray(manifolds, dtype=abi.manifold_dtype)
        c = np.ascontiguousarray(contacts, dtype=abi.contact_dtype)
        self._keep = [m, c]  # the upload is asynchronous
        self.n_contacts = len(c)
        self._chk(self.lib.nb2_upload_manifolds(self.h, abi.ptr(m), ctypes.c_uint32(len(m)), abi.ptr(c),
                                                ctypes.c_uint32(len(c))))

    def upload_manifolds_raw(self, m_ptr, nm, c_ptr, nc):
        """Upload from caller-managed (e.g. pinned) memory; pointers are integers."""
        self.n_contacts = nc
        self._chk(self.lib.nb2_upload_manifolds(self.h, ctypes.c_void_p(m_ptr), ctypes.c_uint32(nm),
                                                ctypes.c_void_p(c_ptr), ctypes.c_uint32(nc)))

    def update_contacts(self, updates):
        """Per-step refresh (world1, world2, normal, depth) of the contacts of the last upload_manifolds."""
        u = np.ascontiguousarray(updates, dtype=abi.contact_update_dtype)
        self._keep_upd = u  # asynchronous
        self._chk(self.lib.nb2_update_contacts(self.h, abi.ptr(u), ctypes.c_uint32(len(u))))

    def update_contacts_raw(self, ptr, n):
        self._chk(self.lib.nb2_update_contacts(self.h, ctypes.c_void_p(ptr), ctypes.c_uint32(n)))

    # ---- device manifold producer
    def upload_colliders(self, colliders):
        c = np.ascontiguousarray(colliders, dtype=abi.collider_dtype)
        self._chk(self.lib.nb2_upload_colliders(self.h, abi.ptr(c), ctypes.c_uint32(len(c))))

    def detect_pairs(self, linear_prediction=0.001, search_radius=-1.0, flip_fraction=0.0):
        n = ctypes.c_uint32()
        self._chk(self.lib.nb2_detect_pairs(self.h, ctypes.c_float(linear_prediction), ctypes.c_float(search_radius),
                                            ctypes.c_uint32(int(flip_fraction * 1000)), ctypes.byref(n)))
        self.n_pairs = int(n.value)
        return self.n_pairs

    def generate_manifolds(self):
        self._chk(self.lib.nb2_generate_manifolds(self.h))
        self.n_contacts = 4 * self.n_pairs

    def download_manifolds(self, compact=False):
        """The contact set the next step will solve, as it sits on the device.  compact=True drops manifolds
        without contacts and closes the gaps of the device producer's 4-slots-per-pair layout, i.e. returns
        the list a host-side producer would have uploaded (plus the slot of every kept contact)."""
        nm, nc = ctypes.c_uint32(), ctypes.c_uint32()
        self._chk(self.lib.nb2_download_manifolds(self.h, None, ctypes.c_uint32(0), None, ctypes.c_uint32(0),
                                                  ctypes.byref(nm), ctypes.byref(nc)))
        m = np.zeros(nm.value, dtype=abi.manifold_dtype)
        c = np.zeros(nc.value, dtype=abi.contact_dtype)
        self._chk(self.lib.nb2_download_manifolds(self.h, abi.ptr(m), nm, abi.ptr(c), nc, ctypes.byref(nm),
                                                  ctypes.byref(nc)))
        if not compact:
            return m, c
        keep_m = m["num_contacts"] > 0
        mm = m[keep_m].copy()
        counts = mm["num_contacts"].astype(np.int64)
        starts = mm["first_contact"].astype(np.int64)
        slots = np.repeat(starts, counts) + (np.arange(int(counts.sum())) - np.repeat(np.cumsum(counts) - counts, counts))
        mm["first_contact"] = np.cumsum(counts) - counts
        return mm, c[slots].copy(), slots

    def upload_joints(self, joints):
        j = np.ascontiguousarray(joints, dtype=abi.joint_dtype)
        self.n_joints = len(j)
        self._chk(self.lib.nb2_upload_joints(self.h, abi.ptr(j), ctypes.c_uint32(len(j))))

    # ---- reduced-coordinate multibodies (Multibody / MultibodyDesc)
    def upload_multibodies(self, multibodies, links):
        m = np.ascontiguousarray(multibodies, dtype=abi.multibody_dtype)
        l = np.ascontiguousarray(links, dtype=abi.mb_link_dtype)
        self.n_mb_links = len(l)
        self._chk(self.lib.nb2_upload_multibodies(self.h, abi.ptr(m), ctypes.c_uint32(len(m)), abi.ptr(l),
                                                  ctypes.c_uint32(len(l))))

    def download_multibody_links(self):
        out = np.zeros(self.n_mb_links, dtype=abi.mb_link_dtype)
        self._chk(self.lib.nb2_download_multibody_links(self.h, abi.ptr(out), ctypes.c_uint32(len(out))))
        return out

    def set_contact_model(self, model):
        """0 = SignoriniCoulombPyramidModel (default), 1 = SignoriniModel (frictionless)."""
        self._chk(self.lib.nb2_set_contact_model(self.h, ctypes.c_int(int(model))))

    def clear_impulse_cache(self):
        self._chk(self.lib.nb2_clear_impulse_cache(self.h))

    # ---- sleeping (ActivationManager)
    def upload_activation(self, activation):
        a = np.ascontiguousarray(activation, dtype=abi.activation_dtype)
        self._chk(self.lib.nb2_upload_activation(self.h, abi.ptr(a), ctypes.c_uint32(len(a))))

    def update_activation(self, mix_factor=0.01, to_activate=()):
        lst = np.ascontiguousarray(to_activate, dtype=np.int32)
        self._chk(self.lib.nb2_update_activation(self.h, ctypes.c_float(mix_factor),
                                                 abi.ptr(lst) if len(lst) else None, ctypes.c_uint32(len(lst))))

    def download_activation(self):
        out = np.zeros(self.n_bodies, dtype=abi.activation_dtype)
        self._chk(self.lib.nb2_download_activation(self.h, abi.ptr(out), ctypes.c_uint32(len(out))))
        return out

    def label_islands(self):
        """(island label per body or -1, velocity rows booked on each body) from the device labelling."""
        lab = np.zeros(self.n_bodies, dtype=np.int32)
        rows = np.zeros(self.n_bodies, dtype=np.uint32)
        self._chk(self.lib.nb2_label_islands(self.h, abi.ptr(lab), abi.ptr(rows), ctypes.c_uint32(self.n_bodies)))
        return lab, rows

    def step(self, mode=abi.MODE_COLOURED):
        self._chk(self.lib.nb2_step(self.h, int(mode)))

    def step_ccd(self, mode=abi.MODE_COLOURED):
        self._chk(self.lib.nb2_step_ccd(self.h, int(mode)))

    def synchronize(self):
        self._chk(self.lib.nb2_synchronize(self.h))

    def download_body_states(self, first=0, n=None, out=None):
        n = self.n_bodies - first if n is None else n
        if out is None:
            out = np.zeros(n, dtype=abi.body_state_dtype)
        self._chk(self.lib.nb2_download_body_states(self.h, abi.ptr(out), ctypes.c_uint32(first),
                                                    ctypes.c_uint32(n)))
        return out

    def download_body_states_raw(self, ptr, first, n):
        self._chk(self.lib.nb2_download_body_states(self.h, ctypes.c_void_p(ptr), ctypes.c_uint32(first),
                                                    ctypes.c_uint32(n)))

    def download_contact_impulses(self):
        out = np.zeros((self.n_contacts, 3), dtype=np.float32)
        if self.n_contacts:
            self._chk(self.lib.nb2_download_contact_impulses(self.h, abi.ptr(out),
                                                             ctypes.c_uint32(self.n_contacts)))
        return out

    def download_joints(self):
        out = np.zeros(self.n_joints, dtype=abi.joint_dtype)
        if self.n_joints:
            self._chk(self.lib.nb2_download_joints(self.h, abi.ptr(out), ctypes.c_uint32(self.n_joints)))
        return out

    def get_stats(self):
        out = np.zeros((), dtype=abi.stats_dtype)
        self._chk(self.lib.nb2_get_stats(self.h, abi.ptr(out)))
        return out

    TIMER_NAMES = ["assembly", "velocity_resolution", "velocity_update", "position_resolution", "step",
                   "velocity_kernel", "position_kernel", "schedule"]

    def get_timers(self):
        """Stage times (ms) of the last step; needs enable_timers()."""
        out = np.zeros(8, dtype=np.float32)
        self._chk(self.lib.nb2_get_timers(self.h, abi.ptr(out)))
        return dict(zip(self.TIMER_NAMES, [float(x) for x in out]))

    def download_schedule(self):
        """(phase, dynamic body of side 1 or -1, of side 2 or -1) per constraint group of the last step."""
        n = ctypes.c_uint32()
        self._chk(self.lib.nb2_download_schedule(self.h, None, None, None, ctypes.c_uint32(0), ctypes.byref(n)))
        ph, a, b = (np.full(n.value, -1, dtype=np.int32) for _ in range(3))
        if n.value:
            self._chk(self.lib.nb2_download_schedule(self.h, abi.ptr(ph), abi.ptr(a), abi.ptr(b), n, ctypes.byref(n)))
        return ph, a, b

    def launch_count(self):
        v = ctypes.c_uint64()
        self._chk(self.lib.nb2_launch_count(self.h, ctypes.byref(v)))
        return int(v.value)
