import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both the oracle and the CUDA library are built in-tree (nvcc cross-compiles without a GPU)."""
    import oracle
    oracle.build()
    from nphysics_b200 import solver
    solver.build()


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_err_q(a, b, floor):
    """Per-quantity relative error (north_star: "within 1e-5 relative (f32)"): every element is compared
    with its OWN reference value, max_i |a_i - b_i| / max(|b_i|, floor).  `floor` is the magnitude below
    which a quantity is judged absolutely (f32 cannot hold 1e-5 relative on a value that is a rounding
    residue of much larger operands): 1e-3 m / rad for poses, 1e-3 m/s for velocities, 1e-6 N s for
    impulses in the tests below."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


class Harness:
    """Drives two solver-like objects (CUDA Solver / Oracle) over the same inputs."""

    def __init__(self, scene, gen=None, params=None):
        self.scene = scene
        self.gen = gen
        self.params = params if params is not None else scene.params

    def setup(self, s):
        s.set_params(self.params)
        s.upload_bodies(self.scene.bodies)
        if len(self.scene.joints):
            s.upload_joints(self.scene.joints)
        return s

    def contacts_for(self, positions):
        from nphysics_b200 import abi
        if self.gen is None:
            return np.zeros(0, abi.manifold_dtype), np.zeros(0, abi.contact_dtype)
        return self.gen.generate(positions)
