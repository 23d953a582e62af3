"""Short K1 invocation for ncu: final scene 1200x675, 60 spp, 3 frames through the C ABI."""
import sys
sys.path.insert(0, '.')
import uecraytracing_b200 as u
exact = 'exact' in sys.argv[1:]
scan = 'tensor' if 'tensor' in sys.argv[1:] else ('fp32' if 'fp32' in sys.argv[1:] else None)
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
for _ in range(3):
    rgb, _, st = u.render(sc, cam, 1200, 60, 50, seed=1, exact=exact, scan=scan)
print(st['kernel_ms'], st['rays_traced'])
