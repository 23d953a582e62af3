"""Small invocation of every kernel (resident and tiled, fast and exact) for compute-sanitizer."""
import sys
import numpy as np
sys.path.insert(0, '.')
import uecraytracing_b200 as u
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
rng = np.random.default_rng(0)
o = rng.uniform(-14, 14, (300, 3)).astype(np.float32); d = rng.normal(size=(300, 3)).astype(np.float32)
for tile in (0, 64):
    u.set_tile_spheres(tile)
    for exact in (False, True):
        t, n, i, f = u.primary_hits(sc, o, d, exact=exact)
        rgb, acc, st = u.render(sc, cam, 48, 3, 50, seed=1, exact=exact, return_accum=True)
        assert np.all(acc[..., 3] == 3), 'sample count'
        print('tile', tile, 'exact', exact, 'hits', int((i >= 0).sum()), 'rays', st['rays_traced'], 'tiles', st['tiles'])
u.set_tile_spheres(0)
print(u.philox(1, [[1, 2, 3, 4]])[0].tolist())
ok, att, oo, do = u.scatter([0, 1, 2], [[.5, .5, .5, 0], [.5, .5, .5, .2], [1, 1, 1, 1.5]], [[0, 0, -1]] * 3, [[0, 0, 0]] * 3, [[0, 0, 1]] * 3, [1, 1, 1], [[.1, .2, .3, .4]] * 3)
print(ok.tolist())
