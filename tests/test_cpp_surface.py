"""The C++ yk:: host surface used as the reference's source.cpp uses its own (tests/cpp/test_yk_surface.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_yk_surface")


@pytest.mark.skipif(not os.path.exists(EXE), reason="build/test_yk_surface not built (make)")
def test_yk_surface_cpu():
    r = subprocess.run([EXE, "--cpu-only"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "cpu checks ok", r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(EXE), reason="build/test_yk_surface not built (make)")
def test_yk_surface_gpu():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.split("\n")[:2] == ["cpu checks ok", "gpu checks ok"], r.stdout + r.stderr


RT = os.path.join(ROOT, "build", "test_yk_raytracer")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RT), reason="build/test_yk_raytracer not built (make)")
def test_yk_raytracer_and_reference_header_names_gpu():
    """source.cpp:100-112 + world.hit / scatter<U> / raytracer::ray_color compiled against the mirror under the reference's
    header names; SURVEY App. B known answers incl. ray_color(centre, mt19937(7)) (tests/cpp/test_yk_raytracer.cpp)."""
    r = subprocess.run([RT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("yk raytracer ok"), r.stdout + r.stderr


TC = os.path.join(ROOT, "build", "test_tc_pack")


@pytest.mark.skipif(not os.path.exists(TC), reason="build/test_tc_pack not built (make)")
def test_tensor_scan_operand_packing_and_bound():
    """Host emulation of what the tensor cores are given (tests/cpp/test_tc_pack.cpp): TF32-exact slots, the UMMA
    canonical layout, and the conservative-superset margin against the reference's double discriminant."""
    r = subprocess.run([TC], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("tc pack ok"), r.stdout + r.stderr


GRID = os.path.join(ROOT, "build", "test_grid_build")


@pytest.mark.skipif(not os.path.exists(GRID), reason="build/test_grid_build not built (make)")
def test_uniform_grid_builder_invariants():
    """Every surface point of every sphere lies in a cell that lists it (tests/cpp/test_grid_build.cpp)."""
    r = subprocess.run([GRID], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("grid build ok"), r.stdout + r.stderr
