"""the UNMODIFIED reference binary (oracle/_ref/raytrace_ref = g++ -std=c++20 -O2
source.cpp, 400x225, 100 spp, depth 50, one core, non-deterministic random_device seeds), timed on this
box, and its PNG compared with the drop-in binary's render of the same (default) configuration."""
import json, os, subprocess, sys, time
import numpy as np
from PIL import Image
sys.path.insert(0, '.')
ref_bin, our_bin = 'oracle/_ref/raytrace_ref', 'uecraytracing_b200/bin/raytrace'
out = {}
def psnr(a, b):
    return float(10 * np.log10(255.0 ** 2 / np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))
imgs = {}
procs = []
for tag in ('ref_a', 'ref_b'):                       # two runs in parallel on two cores: timing + the reference's own noise floor
    procs.append((tag, time.time(), subprocess.Popen([ref_bin, f'gpurun_out/{tag}.png'], stdout=subprocess.PIPE, text=True)))
for tag, t0, p in procs:
    so = p.communicate()[0]; dt = time.time() - t0
    assert p.returncode == 0 and so.split() == ['rendering...', 'rendering', 'finished', 'write', 'to', 'file', ':', f'gpurun_out/{tag}.png', 'success'], so
    out[f'{tag}_seconds'] = dt; out[f'{tag}_samples_per_s_1core'] = 400 * 225 * 100 / dt
    imgs[tag] = np.asarray(Image.open(f'gpurun_out/{tag}.png').convert('RGB'))
t0 = time.time(); r = subprocess.run([our_bin, 'gpurun_out/ours.png'], capture_output=True, text=True); dt = time.time() - t0
assert r.returncode == 0 and r.stdout.split('\n')[:4] == ['rendering...', 'rendering finished', 'write to file : gpurun_out/ours.png', 'success'], r.stdout + r.stderr
imgs['ours'] = np.asarray(Image.open('gpurun_out/ours.png').convert('RGB'))
out['ours_wall_seconds_incl_process_start'] = dt
out['psnr8_ref_a_vs_ref_b'] = psnr(imgs['ref_a'], imgs['ref_b'])
out['psnr8_ours_vs_ref_a'] = psnr(imgs['ours'], imgs['ref_a']); out['psnr8_ours_vs_ref_b'] = psnr(imgs['ours'], imgs['ref_b'])
out['mean8'] = {k: float(v.mean()) for k, v in imgs.items()}
out['shape'] = list(imgs['ours'].shape)
json.dump(out, open('gpurun_out/reference_binary_r1.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
