"""Short swept-tensor-scan K1 invocation for ncu: python tools/profile_sweep.py [spp]"""
import sys
sys.path.insert(0, '.')
import uecraytracing_b200 as u
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc, cam = u.Scene.named('dense', seed=1, n=10000), u.Camera.named('dense')
for _ in range(3):
    rgb, _, st = u.render(sc, cam, 1920, spp, 50, seed=1, scan='tensor')
print(st['kernel_ms'], st['rays_traced'], st['scan_mode'])
