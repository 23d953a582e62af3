"""First-light check on a B200: hit parity, a small render vs the oracle, and a timing run."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import uecraytracing_b200 as u
from oracle import pyoracle

orc = pyoracle.load('best'); print('oracle kind:', orc.kind)

def psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / mse) if mse > 0 else 99.0

def primary_rays(cam, W, H):
    ys, xs = np.mgrid[0:H, 0:W]
    uu = (xs.ravel() + 0.5) / W; vv = (H - ys.ravel() - 1 + 0.5) / H
    c = cam.s
    ll = np.array(c.lower_left, np.float64); hor = np.array(c.horizontal, np.float64); ver = np.array(c.vertical, np.float64); org = np.array(c.origin, np.float64)
    d = ll[None] + uu[:, None] * hor[None] + vv[:, None] * ver[None] - org[None]
    o = np.broadcast_to(org, d.shape)
    return o.astype(np.float32), d.astype(np.float32)

print('philox KAT', [hex(x) for x in u.philox(0, [[0, 0, 0, 0]])[0]])
for name, camname, W in (('ref4', 'ref', 400), ('final', 'final', 600)):
    sc = u.Scene.named(name, seed=1); cam = u.Camera.named(camname); H = u.image_height(W)
    o, d = primary_rays(cam, W, H)
    t0 = time.time(); to, no, io, fo = orc.hits(sc, o.astype(np.float64), d.astype(np.float64)); t1 = time.time()
    for exact in (True, False):
        t, n, i, f = u.primary_hits(sc, o, d, exact=exact)
        hit = io >= 0
        both = hit & (i == io)
        rel = np.abs(t[both] - to[both]) / np.abs(to[both])
        nerr = np.abs(n[both] - no[both]).max(axis=1)
        print(f'{name} exact={exact}: rays {len(o)} hits {hit.sum()} idx mismatch {(i != io).sum()} ff mismatch {(f[both] != fo[both]).sum()} '
              f't rel max {rel.max():.3e} p99 {np.percentile(rel, 99):.3e} normal abs max {nerr.max():.3e} (oracle {t1 - t0:.1f}s)')

# small render vs oracle
sc = u.Scene.named('ref4'); cam = u.Camera.named('ref'); W = 200; H = u.image_height(W); spp = 64
rgb, acc, st = u.render(sc, cam, W, spp, 50, seed=7, return_accum=True)
print('ref4 render stats', {k: st[k] for k in ('rays_traced', 'paths', 'kernel_ms', 'k1_regs', 'k1_blocks', 'k1_smem_bytes')})
t0 = time.time(); osum, orays = orc.render(sc, cam, W, H, spp, 50, 1); t1 = time.time()
o8 = orc.to_color3b(osum, spp)
osum2, _ = orc.render(sc, cam, W, H, spp, 50, 1000003)
print(f'ref4 {W}x{H}x{spp}: PSNR gpu-vs-oracle {psnr8(rgb, o8):.2f} dB, oracle-vs-oracle {psnr8(orc.to_color3b(osum2, spp), o8):.2f} dB; '
      f'rays/sample gpu {st["rays_traced"] / (W * H * spp):.3f} oracle {orays / (W * H * spp):.3f}; mean gpu {acc[..., :3].mean() / spp:.5f} oracle {osum.mean() / spp:.5f}; oracle {t1 - t0:.1f}s')
assert np.all(acc[..., 3] == spp), 'sample count channel'
u.write_png('gpurun_out/first_light_ref4.png', rgb)

sc = u.Scene.named('final', seed=1); cam = u.Camera.named('final'); W = 320; H = u.image_height(W); spp = 32
rgb, acc, st = u.render(sc, cam, W, spp, 50, seed=7, return_accum=True)
t0 = time.time(); osum, orays = orc.render(sc, cam, W, H, spp, 50, 1); t1 = time.time()
o8 = orc.to_color3b(osum, spp); osum2, _ = orc.render(sc, cam, W, H, spp, 50, 1000003)
print(f'final {W}x{H}x{spp}: PSNR gpu-vs-oracle {psnr8(rgb, o8):.2f} dB, oracle-vs-oracle {psnr8(orc.to_color3b(osum2, spp), o8):.2f} dB; '
      f'rays/sample gpu {st["rays_traced"] / (W * H * spp):.3f} oracle {orays / (W * H * spp):.3f}; mean gpu {acc[..., :3].mean() / spp:.5f} oracle {osum.mean() / spp:.5f}; oracle {t1 - t0:.1f}s '
      f'({W * H * spp / (t1 - t0):.3e} samples/s on {__import__("os").cpu_count()} threads)')
u.write_png('gpurun_out/first_light_final.png', rgb)

# timing: final scene 1200x675
for spp in (20, 100):
    for exact in (False, True):
        rgb, _, st = u.render(sc, cam, 1200, spp, 50, seed=3, exact=exact)
        samples = 1200 * 675 * spp
        tf = 18.0 * st['sphere_tests'] / (st['kernel_ms'] * 1e-3) / 1e12
        print(f'final 1200x675x{spp} exact={exact}: K1 {st["kernel_ms"]:.2f} ms  {samples / st["kernel_ms"] * 1e3:.3e} samples/s  rays/sample {st["rays_traced"] / samples:.3f}  '
              f'{tf:.2f} TFLOP/s algorithmic = {tf / 74.45 * 100:.1f}% of FP32 peak; resolve {st["resolve_ms"]:.3f} ms d2h {st["d2h_ms"]:.3f} ms')
u.write_png('gpurun_out/first_light_final_1200.png', rgb)
