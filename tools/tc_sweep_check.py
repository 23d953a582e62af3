"""Swept tensor-core scan (scenes above the residency limit) vs FP32 scan: bit-identity of hits, render agreement, K1 timing.
Run on a B200: python tools/tc_sweep_check.py [spp]"""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import uecraytracing_b200 as u

rng = np.random.default_rng(11)
for n in (1025, 1536, 3000, 10000):
    sc = u.Scene.named('dense', seed=1, n=n)
    o = np.concatenate([rng.uniform(-50, 50, (100000, 1)), rng.uniform(0.05, 6, (100000, 1)), rng.uniform(-50, 50, (100000, 1))], axis=1).astype(np.float32)
    d = (rng.normal(size=(100000, 3)) * rng.uniform(0.05, 5, (100000, 1))).astype(np.float32)
    for exact in (False, True):
        out = {}
        for mode in ("fp32", "tensor"):
            u.set_scan_mode(mode)
            t0 = time.time()
            out[mode] = u.primary_hits(sc, o, d, exact=exact) + (time.time() - t0,)
        u.set_scan_mode("auto")
        a, b = out["fp32"], out["tensor"]
        print(f"n={n} {'exact' if exact else 'fast '}: hits {int((a[2] >= 0).sum())} id mismatches {int((a[2] != b[2]).sum())} t identical {np.array_equal(a[0], b[0])} "
              f"normals identical {np.array_equal(a[1], b[1])} ff {int((a[3] != b[3]).sum())}  ({a[4]:.2f} s / {b[4]:.2f} s)", flush=True)

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc, cam = u.Scene.named('dense', seed=1, n=10000), u.Camera.named('dense')
res = {}
for mode in ("fp32", "tensor", "tensor", "fp32"):
    rgb, acc, st = u.render(sc, cam, 1920, spp, 50, seed=3, return_accum=True, scan=mode)
    n = 1920 * 1080 * spp
    tf = 18.0 * st['sphere_tests'] / (st['kernel_ms'] * 1e-3) / 1e12
    print(f"{mode:6s}: K1 {st['kernel_ms']:.2f} ms {n / st['kernel_ms'] * 1e3:.4e} samples/s {tf:.2f} algorithmic TF = {tf / 74.45 * 100:.1f}% of FP32 peak; "
          f"mode {st['scan_mode']} threads {st['k1_threads']} regs {st['k1_regs']} blocks {st['k1_blocks']} smem {st['k1_smem_bytes']} "
          f"rays {st['rays_traced']} paths {st['paths']}", flush=True)
    res[mode] = (rgb, acc, st)
a, b = res["fp32"][1], res["tensor"][1]
print("accum max abs diff", float(np.abs(a - b).max()), "rgb differing bytes", int((res['fp32'][0] != res['tensor'][0]).sum()),
      "rays equal", res['fp32'][2]['rays_traced'] == res['tensor'][2]['rays_traced'])
