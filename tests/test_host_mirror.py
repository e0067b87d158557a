"""The C++ host mirror (nphysics_b200/host/) compiles against the C ABI; without a GPU it fails
loudly (no fallback); on a B200 it runs examples3d/pyramid3 through MechanicalWorld::step."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "nphysics_b200", "host")
EXE = os.path.join(HOST, "example_pyramid3")


def build_example():
    src = os.path.join(HOST, "example_pyramid3.cpp")
    hdr = os.path.join(HOST, "nphysics_b200.hpp")
    if (not os.path.exists(EXE)) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", src, "-o", EXE, "-L" + os.path.join(ROOT, "nphysics_b200"),
                               "-lnphysics_b200", "-Wl,-rpath,$ORIGIN/.."])
    return EXE


def test_host_mirror_compiles_and_links():
    exe = build_example()
    assert os.path.exists(exe)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_host_mirror_has_no_cpu_fallback():
    exe = build_example()
    r = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert r.returncode == 2
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["coloured", "reference"])
def test_pyramid3_example_runs(mode):
    exe = build_example()
    r = subprocess.run([exe, "40", mode], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "initial manifolds 1335 contacts 5340 rows 16020" in r.stdout
    assert r.stdout.strip().endswith("OK")
