"""north_star gate on the SHIPPED default kernel: RTIOW final scene 1200x675, 500 spp, depth 50, PSNR of the GPU frame against the two
independent 500-spp renders by the unmodified reference code and their mean (tests/golden/ref_final_1200x675*.png).
Run on a B200: python tools/psnr_gate.py > gpurun_out/psnr_gate_r2.json"""
import json, os, sys, zlib
import numpy as np
sys.path.insert(0, '.')
import uecraytracing_b200 as u

def read_png(path):
    raw = open(path, 'rb').read(); w, h = int.from_bytes(raw[16:20], 'big'), int.from_bytes(raw[20:24], 'big')
    pos, idat = 8, b''
    while pos < len(raw):
        n = int.from_bytes(raw[pos:pos + 4], 'big')
        if raw[pos + 4:pos + 8] == b'IDAT': idat += raw[pos + 8:pos + 8 + n]
        pos += 12 + n
    return np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)[:, 1:].reshape(h, w, 3)

def psnr(a, b):
    return float(10 * np.log10(255.0 ** 2 / np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))

G = 'tests/golden'
meta = json.load(open(f'{G}/ref_final_1200x675.json'))
out = {k: meta[k] for k in ('scene', 'width', 'height', 'spp_each', 'depth', 'oracle', 'psnr8_A_vs_B', 'linear_mean', 'rays_per_sample_A') if k in meta}
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
for label, kw in (('default', {}), ('exact', {'exact': True}), ('fp32_scan', {'scan': 'fp32'})):
    rgb, acc, st = u.render(sc, cam, 1200, 500, 50, seed=2026, return_accum=True, **kw)
    r = {'scan_mode': st['scan_mode'], 'k1_threads': st['k1_threads'], 'kernel_ms': st['kernel_ms'], 'rays_per_sample': st['rays_traced'] / st['paths'],
         'linear_mean': float(acc[..., :3].mean() / 500), 'psnr8_vs_mean_1000spp': psnr(rgb, read_png(f'{G}/ref_final_1200x675_1000spp.png'))}
    for s in ('A', 'B'):
        p = f'{G}/ref_final_1200x675_{s}.png'
        if os.path.exists(p): r[f'psnr8_vs_{s}'] = psnr(rgb, read_png(p))
    out[f'gpu500_{label}'] = r
out['note'] = ('default = automatic scan (tensor cores, self-scan pipeline k_render_paths_ts, k1_threads 896); gate: PSNR >= 40 dB; the noise floor of any 500-spp '
               'render against another is psnr8_A_vs_B, against the 1000-spp mean floor + 1.25 dB')
print(json.dumps(out, indent=1))
