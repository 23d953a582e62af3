"""K1 time of a short launch (the per-GPU share of an 8-GPU 500-spp frame) against a long one: the fixed start-up + drain cost.
UECRT_LIB selects a build: python tools/tail_test.py"""
import sys; sys.path.insert(0,'.')
import uecraytracing_b200 as u
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
res = {}
for spp in (63, 500, 63, 500, 63, 500):
    rgb, acc, st = u.render(sc, cam, 1200, spp, 50, seed=3, return_accum=True, scan='tensor')
    res.setdefault(spp, []).append(st['kernel_ms'])
a, b = min(res[63]), min(res[500])
slope = (b - a) / (500 - 63)
print(f"63 spp {a:.3f} ms, 500 spp {b:.3f} ms, per spp {slope * 1e3:.2f} us, fixed cost {a - 63 * slope:.3f} ms", flush=True)
