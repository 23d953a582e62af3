"""FAST (FP32) resolve against the double oracle on primary and random rays: t and normal error percentiles.
Run on a B200: python tools/fast_accuracy.py"""
import sys
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import uecraytracing_b200 as u
from conftest import primary_rays
from oracle import pyoracle

orc = pyoracle.load('best')
for name, cam, W in (("ref4", "ref", 400), ("final", "final", 1200)):
    sc, c = u.Scene.named(name, seed=1), u.Camera.named(cam)
    o, d = primary_rays(c, W, u.image_height(W))
    rng = np.random.default_rng(3)
    o2 = rng.uniform(-14, 14, (300000, 3)).astype(np.float32)
    d2 = (rng.normal(size=(300000, 3)) * rng.uniform(0.05, 5, (300000, 1))).astype(np.float32)
    for label, oo, dd in (("primary", o, d), ("random", o2, d2)):
        to, no, io, fo = orc.hits(sc, oo.astype(np.float64), dd.astype(np.float64))
        for mode in ("fp32", "tensor"):
            u.set_scan_mode(mode)
            t, n, i, f = u.primary_hits(sc, oo, dd, exact=False)
            u.set_scan_mode("auto")
            both = (io >= 0) & (i == io)
            rel = np.abs(t[both] - to[both]) / np.abs(to[both])
            dn = np.abs(n[both] - no[both]).max(axis=1)
            print(f"{name:6s} {label:8s} {mode:6s}: idx mismatches {int((i != io).sum())} / {len(i)}; t rel p50 {np.percentile(rel, 50):.2e} p99 {np.percentile(rel, 99):.2e} "
                  f"max {rel.max():.2e}; normal abs p99 {np.percentile(dn, 99):.2e} max {dn.max():.2e}; front_face mismatches {int((f[both] != fo[both]).sum())}", flush=True)
            if mode == "fp32":
                idx = np.flatnonzero(both)[np.argsort(-rel)[:3]]
                for k in idx:
                    if abs(t[k] - to[k]) > 1e-6 * abs(to[k]):
                        g = sc.geom[io[k]]
                        oc = oo[k].astype(np.float64) - g[:3]; D = dd[k].astype(np.float64)
                        a = D @ D; hb = oc @ D; cc = oc @ oc - g[3] ** 2; disc = hb * hb - a * cc
                        print(f"   ray {k}: id {i[k]} sphere {g} t {t[k]!r} oracle {to[k]!r} ff {f[k]} / {fo[k]} disc {disc:.6e} roots {(-hb - np.sqrt(max(disc, 0))) / a!r} {(-hb + np.sqrt(max(disc, 0))) / a!r} |d| {np.sqrt(a)!r}")
