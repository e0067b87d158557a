"""The C++ host mirror (nphysics_b200/host/) compiles against the C ABI; without a GPU it fails
loudly (no fallback); on a B200 it runs examples3d/pyramid3 through MechanicalWorld::step."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "nphysics_b200", "host")
EXE = os.path.join(HOST, "example_pyramid3")


def build_example(name="example_pyramid3"):
    exe = os.path.join(HOST, name)
    src = os.path.join(HOST, name + ".cpp")
    hdr = os.path.join(HOST, "nphysics_b200.hpp")
    if (not os.path.exists(exe)) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", src, "-o", exe, "-L" + os.path.join(ROOT, "nphysics_b200"),
                               "-lnphysics_b200", "-Wl,-rpath,$ORIGIN/.."])
    return exe


def test_host_mirror_compiles_and_links():
    for name in ("example_pyramid3", "example_ragdoll3", "example_ragdoll3_colliders"):
        assert os.path.exists(build_example(name))


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


@pytest.mark.gpu
def test_ragdoll3_example_runs():
    """examples3d/ragdoll3.rs as shipped (one Multibody per ragdoll) through MultibodyDesc / MechanicalWorld::step of the
    C++ mirror: the program itself checks that no joint anchor drifts."""
    exe = build_example("example_ragdoll3")
    r = subprocess.run([exe, "120"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "27 multibodies (162 links)" in r.stdout


@pytest.mark.gpu
def test_ragdoll3_with_colliders_example_runs():
    """The reference's main loop shape -- `mechanical_world.step(geometrical_world, bodies, colliders, joints)` -- over
    the C++ mirror: ground collider + one cuboid collider per link in a DefaultColliderSet, contacts produced on the
    device every step.  The program checks joint drift, that nothing falls through the ground, finiteness."""
    exe = build_example("example_ragdoll3_colliders")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "8 multibodies (48 links, 49 colliders" in r.stdout
