"""Short grid-walk K1 invocation for ncu: python tools/profile_grid.py [final|dense]"""
import sys
sys.path.insert(0, '.')
import uecraytracing_b200 as u
which = sys.argv[1] if len(sys.argv) > 1 else 'dense'
sc, cam = (u.Scene.named('dense', seed=1, n=10000), u.Camera.named('dense')) if which == 'dense' else (u.Scene.named('final', seed=1), u.Camera.named('final'))
W, spp = (1920, 16) if which == 'dense' else (1200, 60)
for _ in range(3):
    rgb, _, st = u.render(sc, cam, W, spp, 50, seed=1, scan='grid')
print(st['kernel_ms'], st['rays_traced'], st['sphere_tests'])
