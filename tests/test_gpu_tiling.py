"""-m gpu: shared-memory tiling of the scene (the two-buffer TMA ring).  Tiling must not change any
result: same candidates, same resolve, same tie rule across tile boundaries."""
import numpy as np
import pytest

from conftest import primary_rays

pytestmark = pytest.mark.gpu


@pytest.fixture
def tiles(u):
    yield u.set_tile_spheres
    u.set_tile_spheres(0)


@pytest.mark.parametrize("tile", [16, 64, 200])
def test_forced_tiles_hits_exact(u, oracle, tiles, tile):
    sc, c = u.Scene.named("final", seed=1), u.Camera.named("final")
    o, d = primary_rays(c, 400, 225)
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    tiles(tile)
    t, n, i, f = u.primary_hits(sc, o, d, exact=True)
    assert np.array_equal(i, want[2]) and np.array_equal(f, want[3])
    hit = want[2] >= 0
    assert np.max(np.abs(t[hit] - want[0][hit]) / want[0][hit]) <= 1e-9
    t2, n2, i2, f2 = u.primary_hits(sc, o[:77], d[:77], exact=False)          # ragged last batch
    assert (i2 != want[2][:77]).sum() == 0


def test_tie_rule_across_tile_boundary(u, tiles):
    """40 coincident spheres cut into tiles of 16: the LAST index must still win (hittable_list.hpp:40-41)."""
    g = np.tile(np.array([[0, 0, -1, .5]], np.float32), (40, 1))
    sc = u.Scene(g, np.zeros(40, np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (40, 1)))
    tiles(16)
    for exact in (True, False):
        t, n, i, f = u.primary_hits(sc, [[0, 0, 0]] * 5, [[0, 0, -1]] * 5, exact=exact)
        assert i.tolist() == [39] * 5 and np.allclose(t, 0.5)


def test_forced_tiles_render_equals_resident(u, tiles):
    sc, c = u.Scene.named("final", seed=1), u.Camera.named("final")
    _, a, sa = u.render(sc, c, 160, 16, 50, seed=21, return_accum=True)
    assert sa["tiles"] == 1
    tiles(64)
    _, b, sb = u.render(sc, c, 160, 16, 50, seed=21, return_accum=True)
    assert sb["tiles"] == 8 and sb["tile_spheres"] == 64
    assert sb["rays_traced"] == sa["rays_traced"]
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-5)      # same paths; only the atomic add order differs
    _, e, _ = u.render(sc, c, 160, 16, 50, seed=21, return_accum=True, exact=True)
    assert np.isclose(a[..., :3], e[..., :3], rtol=1e-3, atol=1e-3).all(axis=-1).mean() > 0.9


def test_auto_tiling_large_scene(u, oracle):
    """20 000 spheres = 320 KB packed -> automatic tiles of 2048 spheres."""
    sc, c = u.Scene.named("dense", seed=2, n=20000), u.Camera.named("dense")
    o, d = primary_rays(c, 320, 180)
    pick = np.random.default_rng(1).choice(len(o), 20000, replace=False)
    o, d = o[pick], d[pick]
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    t, n, i, f = u.primary_hits(sc, o, d, exact=True)
    assert np.array_equal(i, want[2]) and np.array_equal(f, want[3])
    _, acc, st = u.render(sc, c, 192, 4, 50, seed=1, return_accum=True)
    assert st["tiles"] == 10 and st["tile_spheres"] == 2048 and np.all(acc[..., 3] == 4)
    assert st["sphere_tests"] == st["rays_traced"] * 20000
