"""Quick correctness + K1 timing probe used while tuning (final scene, 1200x675)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import uecraytracing_b200 as u
from oracle import pyoracle
orc = pyoracle.load('best')
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
W, H = 400, 225
ys, xs = np.mgrid[0:H, 0:W]; c = cam.s
ll, hor, ver, org = (np.array(getattr(c, k), np.float64) for k in ("lower_left", "horizontal", "vertical", "origin"))
d = (ll[None] + ((xs.ravel() + .5) / W)[:, None] * hor[None] + ((H - ys.ravel() - .5) / H)[:, None] * ver[None] - org[None]).astype(np.float32)
o = np.broadcast_to(org.astype(np.float32), d.shape).copy()
to, no, io, fo = orc.hits(sc, o.astype(np.float64), d.astype(np.float64))
for exact in (True, False):
    t, n, i, f = u.primary_hits(sc, o, d, exact=exact)
    print('exact' if exact else 'fast ', 'idx mismatches', int((i != io).sum()), 'of', len(i))
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 200
import os
if os.environ.get("UECRT_RAYS"): u.set_rays_per_thread(int(os.environ["UECRT_RAYS"]))
for exact in (False, False, True):
    rgb, acc, st = u.render(sc, cam, 1200, spp, 50, seed=3, exact=exact, return_accum=True)
    n = 1200 * 675 * spp
    tf = 18.0 * st['sphere_tests'] / (st['kernel_ms'] * 1e-3) / 1e12
    print(f"exact={exact}: K1 {st['kernel_ms']:.2f} ms {n / st['kernel_ms'] * 1e3:.4e} samples/s {tf:.2f} TF = {tf / 74.45 * 100:.1f}% peak; regs {st['k1_regs']} blocks {st['k1_blocks']} mean {acc[..., :3].mean() / spp:.5f} rays/sample {st['rays_traced'] / n:.4f}")
