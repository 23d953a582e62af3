"""K1 time of the three searches on prefixes of the final scene (where does the automatic policy switch?): python tools/crossover.py"""
import sys
import numpy as np
sys.path.insert(0, '.')
import uecraytracing_b200 as u
full, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
for n in (8, 32, 48, 64, 80, 96, 128, 256, 484):
    sc = u.Scene(full.geom[:n], full.mat_type[:n], full.mat_param[:n])
    row = []
    for mode in ("fp32", "tensor"):
        ts = []
        for _ in range(3):
            _, _, st = u.render(sc, cam, 1200, 20, 50, seed=3, scan=mode)
            ts.append(st['kernel_ms'])
        row.append(min(ts[1:]))
    print(f"n={n:4d}  fp32 {row[0]:.2f} ms  tensor {row[1]:.2f} ms  -> {'tensor' if row[1] < row[0] else 'fp32'}", flush=True)
