"""Full-resolution CPU reference render of BASELINE configs[1] (final scene, 1200x675, depth 50) with
the UNMODIFIED reference headers (oracle/_ref/liboracle_ref.so), for the north_star PSNR gate.

Runs where many host cores are available (the GPU box; ~3.5 min per 500 spp on 16 threads).  Writes
  gpurun_out/ref_final_1200_A.png / _B.png : two independent 500-spp renders (8-bit, to_color3b)
  gpurun_out/ref_final_1200_mean.png       : to_color3b of their average (1000 spp)
  gpurun_out/ref_final_1200.json           : seeds, timing, oracle-vs-oracle PSNR floor, linear stats
and, when a GPU is present, the GPU-vs-reference PSNR of a 500-spp GPU render.
Copy the PNG/JSON you want pinned into tests/golden/ (committed: the GPU box has no /root/reference).
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, '.')
import uecraytracing_b200 as u
from oracle import pyoracle

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 500
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
orc = pyoracle.load('ref')
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
H = u.image_height(W)

def psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float(10 * np.log10(255.0 ** 2 / mse))

out = {'scene': 'final', 'scene_seed': 1, 'width': W, 'height': H, 'spp_each': spp, 'depth': 50, 'oracle': orc.kind, 'threads': os.cpu_count()}
imgs = {}
for tag, seed in (('A', 1), ('B', 2000003)):
    t0 = time.time(); s, rays = orc.render(sc, cam, W, H, spp, 50, seed); dt = time.time() - t0
    imgs[tag] = s
    out[f'seed_{tag}'] = seed; out[f'seconds_{tag}'] = dt; out[f'samples_per_s_{tag}'] = W * H * spp / dt; out[f'rays_per_sample_{tag}'] = rays / (W * H * spp)
    u.write_png(f'gpurun_out/ref_final_{W}_{tag}.png', orc.to_color3b(s, spp))
    print(tag, f'{dt:.1f}s', flush=True)
A8, B8 = orc.to_color3b(imgs['A'], spp), orc.to_color3b(imgs['B'], spp)
mean = (imgs['A'] + imgs['B']) / 2
M8 = orc.to_color3b(mean, spp)
u.write_png(f'gpurun_out/ref_final_{W}_mean.png', M8)
out['psnr8_A_vs_B'] = psnr8(A8, B8)
out['mse_linear_A_vs_B'] = float(np.mean(((imgs['A'] - imgs['B']) / spp) ** 2))
out['linear_mean'] = float(mean.mean() / spp)
if u.device_count() > 0:
    for gspp in (spp,):
        rgb, acc, st = u.render(sc, cam, W, gspp, 50, seed=2026, return_accum=True)
        lin = acc[..., :3] / gspp
        out[f'gpu{gspp}_psnr8_vs_A'] = psnr8(rgb, A8); out[f'gpu{gspp}_psnr8_vs_B'] = psnr8(rgb, B8); out[f'gpu{gspp}_psnr8_vs_mean'] = psnr8(rgb, M8)
        out[f'gpu{gspp}_mse_linear_vs_mean'] = float(np.mean((lin - mean / spp) ** 2))
        out[f'gpu{gspp}_linear_mean'] = float(lin.mean()); out[f'gpu{gspp}_kernel_ms'] = st['kernel_ms']
        u.write_png(f'gpurun_out/gpu_final_{W}_{gspp}spp.png', rgb)
json.dump(out, open(f'gpurun_out/ref_final_{W}.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
