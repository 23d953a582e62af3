"""Tensor-core scan vs FP32 scan: hit parity on identical rays, render agreement, K1 timing.
Run on a B200: python tools/tc_check.py [spp]"""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import uecraytracing_b200 as u
from oracle import pyoracle

orc = pyoracle.load('best')
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
W, H = 400, 225
ys, xs = np.mgrid[0:H, 0:W]
c = cam.s
ll, hor, ver, org = (np.array(getattr(c, k), np.float64) for k in ("lower_left", "horizontal", "vertical", "origin"))
d = (ll[None] + ((xs.ravel() + .5) / W)[:, None] * hor[None] + ((H - ys.ravel() - .5) / H)[:, None] * ver[None] - org[None]).astype(np.float32)
o = np.broadcast_to(org.astype(np.float32), d.shape).copy()
to, no, io, fo = orc.hits(sc, o.astype(np.float64), d.astype(np.float64))
rng = np.random.default_rng(11)
o2 = rng.uniform(-14, 14, (200000, 3)).astype(np.float32)
d2 = (rng.normal(size=(200000, 3)) * rng.uniform(0.05, 5, (200000, 1))).astype(np.float32)
t2o, n2o, i2o, f2o = orc.hits(sc, o2.astype(np.float64), d2.astype(np.float64))
print("tensor scan max spheres", u.tensor_scan_max_spheres(), flush=True)
for mode in ("fp32", "tensor"):
    u.set_scan_mode(mode)
    for exact in (True, False):
        t0 = time.time()
        t, n, i, f = u.primary_hits(sc, o, d, exact=exact)
        t2, n2, i2, f2 = u.primary_hits(sc, o2, d2, exact=exact)
        hit = io >= 0
        rel = np.abs(t[hit] - to[hit]) / np.abs(to[hit])
        print(f"{mode:6s} {'exact' if exact else 'fast '}: primary idx mismatches {int((i != io).sum())} / {len(i)}, t rel max {rel.max():.2e};"
              f" random idx mismatches {int((i2 != i2o).sum())} / {len(i2)}, ff {int((f2 != f2o).sum())}  ({time.time() - t0:.2f} s)", flush=True)
u.set_scan_mode("auto")

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
res = {}
for mode in ("fp32", "tensor", "tensor", "fp32"):
    rgb, acc, st = u.render(sc, cam, 1200, spp, 50, seed=3, return_accum=True, scan=mode)
    n = 1200 * 675 * spp
    tf = 18.0 * st['sphere_tests'] / (st['kernel_ms'] * 1e-3) / 1e12
    print(f"{mode:6s}: K1 {st['kernel_ms']:.2f} ms {n / st['kernel_ms'] * 1e3:.4e} samples/s {tf:.2f} algorithmic TF = {tf / 74.45 * 100:.1f}% of FP32 peak; "
          f"mode {st['scan_mode']} threads {st['k1_threads']} regs {st['k1_regs']} blocks {st['k1_blocks']} smem {st['k1_smem_bytes']} "
          f"mean {acc[..., :3].mean() / spp:.6f} rays {st['rays_traced']} paths {st['paths']}", flush=True)
    res[mode] = (rgb, acc, st)
a, b = res["fp32"][1], res["tensor"][1]
print("accum max abs diff", float(np.abs(a - b).max()), "rgb differing bytes", int((res['fp32'][0] != res['tensor'][0]).sum()),
      "rays equal", res['fp32'][2]['rays_traced'] == res['tensor'][2]['rays_traced'])
rgb, acc, st = u.render(sc, cam, 400, 16, 50, seed=3, return_accum=True, scan="tensor", exact=True)
rgb2, acc2, st2 = u.render(sc, cam, 400, 16, 50, seed=3, return_accum=True, scan="fp32", exact=True)
print("exact: tensor K1", st['kernel_ms'], "fp32 K1", st2['kernel_ms'], "accum max abs diff", float(np.abs(acc - acc2).max()), "rays equal", st['rays_traced'] == st2['rays_traced'])
