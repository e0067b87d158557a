"""The C-ABI library loads and exports every symbol include/nphysics_b200.h declares; host-only
entry points behave; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from nphysics_b200 import abi, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "nphysics_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(?:int|const char\*)\s+(nb2_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_and_binding_agree_on_the_export_list():
    assert declared_functions() == sorted(solver.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(solver.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_struct_sizes_match_the_numpy_mirror():
    lib = solver.load()
    for i, d in enumerate(abi.SIZEOF_ORDER):
        assert lib.nb2_sizeof(i) == d.itemsize
    assert lib.nb2_sizeof(99) == abi.ERR_INVALID_ARGUMENT
    assert lib.nb2_abi_version() == 1


def test_default_params_are_integration_parameters_default():
    """src/solver/integration_parameters.rs:169-189."""
    lib = solver.load()
    p = np.zeros((), abi.params_dtype)
    assert lib.nb2_default_params(abi.ptr(p)) == 0
    q = abi.default_params()
    for name in abi.params_dtype.names:
        assert np.all(p[name] == q[name]), name
    assert p["max_velocity_iterations"] == 8 and p["max_position_iterations"] == 3
    assert p["dt"] == np.float32(1.0 / 60.0) and p["erp"] == np.float32(0.2)


def test_material_combine_precedence():
    """MaterialCombineMode::combine, src/material/material.rs:72-86: Max > Multiply > Min > Average;
    surface velocity is props1 - props2 (:174)."""
    lib = solver.load()
    f = ctypes.c_float
    AVG, MIN, MUL, MAX = 0, 1, 2, 3

    def combine(f1, m1, f2, m2):
        of, orr = f(), f()
        sv = np.zeros(3, np.float32)
        s1 = np.array([1.0, 2.0, 3.0], np.float32)
        s2 = np.array([0.5, 0.0, -1.0], np.float32)
        rc = lib.nb2_combine_materials(f(f1), m1, f(0.1), AVG, abi.ptr(s1), f(f2), m2, f(0.3), AVG, abi.ptr(s2),
                                       ctypes.byref(of), ctypes.byref(orr), abi.ptr(sv))
        assert rc == 0
        assert orr.value == pytest.approx(0.2)
        assert np.allclose(sv, [0.5, 2.0, 4.0])
        return of.value

    assert combine(0.5, AVG, 0.3, AVG) == pytest.approx(0.4)
    assert combine(0.5, MIN, 0.3, AVG) == pytest.approx(0.3)
    assert combine(0.5, MIN, 0.3, MUL) == pytest.approx(0.15)
    assert combine(0.5, MAX, 0.3, MUL) == pytest.approx(0.5)
    assert combine(0.5, AVG, 0.3, MAX) == pytest.approx(0.5)


def test_error_strings():
    lib = solver.load()
    assert lib.nb2_error_string(0) == b"ok"
    assert b"no CPU fallback" in lib.nb2_error_string(abi.ERR_NO_DEVICE)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_create_fails_loudly_without_a_gpu():
    """The product path never falls back to a CPU implementation."""
    with pytest.raises(solver.Nb2Error) as ei:
        solver.Solver(0)
    assert ei.value.code == abi.ERR_NO_DEVICE
    lib = solver.load()
    assert lib.nb2_step(None, 0) == abi.ERR_INVALID_ARGUMENT


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under nphysics_b200/ may reference it."""
    pkg = os.path.join(ROOT, "nphysics_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn
                assert "liboracle" not in text and "nbo_" not in text, fn
