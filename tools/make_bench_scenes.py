"""Writes tests/golden/bench_scenes.npz: the scene and camera arrays of bench.py's workloads, produced by the
product's host scene builders (host/yk/scenes.hpp via libuecrt.so).  bench.py's reference arm and cpu_baseline leg
read this fixture through oracle/pyoracle.py, so the CPU arm never loads libuecrt.so; tests/test_host_logic.py
checks the fixture against the builders."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uecraytracing_b200 as u  # noqa: E402

out = {}
for name in ("final", "ref4", "dense"):
    sc = u.Scene.named(name, seed=1, n=10000)
    out[f"{name}_geom"], out[f"{name}_mat_type"], out[f"{name}_mat_param"] = sc.geom, sc.mat_type, sc.mat_param
for cam in ("final", "ref", "dense"):
    s = u.Camera.named(cam).s
    out[f"cam_{cam}_vec"] = np.array([list(getattr(s, k)) for k in ("origin", "lower_left", "horizontal", "vertical", "u", "v")], np.float32)
    out[f"cam_{cam}_lens"] = np.float32(s.lens_radius)
path = os.path.join(ROOT, "tests", "golden", "bench_scenes.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
