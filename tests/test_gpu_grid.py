"""-m gpu: the opt-in uniform grid (csrc/grid_build.hpp, grid_walk.cuh) behind the same
closest-hit contract as hittable_list::hit_impl (hittable_list.hpp:32-58).  The grid only selects which
spheres a ray tests; the per-sphere test and the tie rule are the brute-force path's, so in the
FP64-resolved mode the result must equal the oracle's bit for bit, and whole renders must trace
exactly the same rays as the brute-force scans."""
import numpy as np
import pytest

from conftest import golden, primary_rays, unhex
from test_gpu_hits import T_REL_EXACT, _check_exact

pytestmark = pytest.mark.gpu


@pytest.fixture()
def grid(u):
    u.set_scan_mode("grid")
    yield u
    u.set_scan_mode("auto")


@pytest.mark.parametrize("name,cam,W", [("ref4", "ref", 400), ("ref4d", "ref", 400), ("final", "final", 1200)])
def test_primary_rays_exact(grid, oracle, name, cam, W):
    u = grid
    sc, c = u.Scene.named(name, seed=1), u.Camera.named(cam)
    o, d = primary_rays(c, W, u.image_height(W))
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    got = u.primary_hits(sc, o, d, exact=True)
    _check_exact(got, want, name + " grid")
    fast = u.primary_hits(sc, o, d, exact=False)
    u.set_scan_mode("fp32")
    for x, y in zip(fast, u.primary_hits(sc, o, d, exact=False)):
        assert np.array_equal(x, y), name + ": FP32 resolve differs between the grid walk and the brute-force scan"


def test_dense_10k_exact(grid, oracle):
    """BASELINE config 3's scene: 10k spheres, sampled primary rays and bounce-like rays from the hit points."""
    u = grid
    sc, c = u.Scene.named("dense", seed=1, n=10000), u.Camera.named("dense")
    o, d = primary_rays(c, 1920, 1080)
    pick = np.random.default_rng(3).choice(len(o), 60000, replace=False)
    o, d = o[pick], d[pick]
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    _check_exact(u.primary_hits(sc, o, d, exact=True), want, "dense grid")
    t, n, i, f = want
    hit = i >= 0
    p = (o[hit].astype(np.float64) + d[hit].astype(np.float64) * t[hit][:, None]).astype(np.float32)
    rng = np.random.default_rng(8)
    v = rng.normal(size=p.shape); v /= np.linalg.norm(v, axis=1, keepdims=True)
    d2 = (n[hit] + v).astype(np.float32)
    _check_exact(u.primary_hits(sc, p, d2, exact=True), oracle.hits(sc, p.astype(np.float64), d2.astype(np.float64)), "dense secondary grid")


@pytest.mark.parametrize("name,span", [("ref4", 3.0), ("final", 14.0)])
def test_random_and_golden_rays_exact(grid, oracle, name, span):
    u = grid
    sc = u.Scene.named(name, seed=1)
    g = golden("random_ray_hits.json")[name]
    o = np.array(g["o"], np.float32).reshape(-1, 3); d = np.array(g["d"], np.float32).reshape(-1, 3)
    want = (unhex(g["t"]), unhex(g["n"], (-1, 3)), np.array(g["id"], np.int32), np.array(g["ff"], np.uint8))
    _check_exact(u.primary_hits(sc, o, d, exact=True), want, name + " golden grid")
    rng = np.random.default_rng(13)
    o = rng.uniform(-span, span, (200000, 3)).astype(np.float32)
    d = (rng.normal(size=(200000, 3)) * rng.uniform(0.05, 5, (200000, 1))).astype(np.float32)
    _check_exact(u.primary_hits(sc, o, d, exact=True), oracle.hits(sc, o.astype(np.float64), d.astype(np.float64)), name + " grid")
    # axis-parallel rays (zero direction components: the DDA's degenerate axes) through the scene
    axes = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 50000)] * rng.choice([-1.0, 1.0], (50000, 1)).astype(np.float32)
    o = rng.uniform(-span, span, (50000, 3)).astype(np.float32)
    _check_exact(u.primary_hits(sc, o, axes, exact=True), oracle.hits(sc, o.astype(np.float64), axes.astype(np.float64)), name + " axis rays")


@pytest.mark.parametrize("shift", [(300.0, 0.0, 300.0), (3000.0, 50.0, -2000.0)])
def test_far_from_origin_stays_exact(grid, oracle, shift):
    u = grid
    base, c = u.Scene.named("final", seed=1), u.Camera.named("final")
    g = base.geom.copy(); g[:, :3] += np.array(shift, np.float32)
    sc = u.Scene(g, base.mat_type, base.mat_param)
    o, d = primary_rays(c, 240, 135)
    o = o + np.array(shift, np.float32)
    _check_exact(u.primary_hits(sc, o, d, exact=True), oracle.hits(sc, o.astype(np.float64), d.astype(np.float64)), f"shift {shift} grid")


def test_edge_cases_ties_and_intervals(grid, oracle):
    u = grid
    for c in golden("hit_edge_cases.json")["cases"]:
        g = np.array(c["spheres"], np.float32)
        sc = u.Scene(g, np.zeros(len(g), np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (len(g), 1)))
        tmax = float("inf") if c["tmax"] == "inf" else float.fromhex(c["tmax"])
        t, n, i, f = u.primary_hits(sc, [c["o"]], [c["d"]], float.fromhex(c["tmin"]), tmax, exact=True)
        assert int(i[0]) == c["id"], c["name"]
        if c["id"] >= 0:
            assert int(f[0]) == c["ff"], c["name"]
            assert abs(t[0] - unhex(c["t"])[0]) <= T_REL_EXACT * abs(unhex(c["t"])[0]), c["name"]
    # coincident spheres spread over the index range: the LAST one wins although the walk meets them in cell order
    n = 300
    rng = np.random.default_rng(2)
    geom = np.concatenate([rng.uniform(-5, 5, (n, 3)), np.full((n, 1), 0.3)], axis=1).astype(np.float32)
    geom[:, 2] -= 20.0
    for k in (3, 40, 131, 299):
        geom[k] = (0, 0, -1, 0.5)
    sc = u.Scene(geom, np.zeros(n, np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (n, 1)))
    for exact in (True, False):
        t, nn, i, f = u.primary_hits(sc, [[0, 0, 0]], [[0, 0, -1]], exact=exact)
        assert int(i[0]) == 299 and abs(t[0] - 0.5) < 1e-6
    empty = u.Scene(np.zeros((0, 4)), np.zeros(0), np.zeros((0, 4)))
    t, nn, i, f = u.primary_hits(empty, [[0, 0, 0]], [[0, 0, -1]])
    assert i[0] == -1
    sc = u.Scene.named("final", seed=1)
    o, d = primary_rays(u.Camera.named("final"), 200, 112)
    for tmin, tmax in ((0.001, 1.2), (1.0, 1.5), (0.5, float("inf"))):
        _check_exact(u.primary_hits(sc, o, d, tmin, tmax), oracle.hits(sc, o.astype(np.float64), d.astype(np.float64), tmin, tmax), f"[{tmin},{tmax}] grid")


@pytest.mark.parametrize("name,cam,W,spp,other", [("ref4d", "ref", 400, 16, "fp32"), ("final", "final", 400, 16, "tensor"), ("dense", "dense", 320, 4, "fp32")])
@pytest.mark.parametrize("exact", [False, True])
def test_render_traces_the_same_paths_as_brute_force(u, name, cam, W, spp, other, exact):
    sc, c = u.Scene.named(name, seed=1, n=3000), u.Camera.named(cam)
    ra, aa, sa = u.render(sc, c, W, spp, 50, seed=9, exact=exact, return_accum=True, scan="grid")
    rb, ab, sb = u.render(sc, c, W, spp, 50, seed=9, exact=exact, return_accum=True, scan=other)
    assert sa["scan_mode"] == 3
    assert sa["rays_traced"] == sb["rays_traced"] and sa["paths"] == sb["paths"] == W * u.image_height(W) * spp
    assert np.array_equal(aa[..., 3], ab[..., 3])
    assert np.abs(aa[..., :3] - ab[..., :3]).max() <= 1e-4 * spp
    assert 0 < sa["sphere_tests"] < sb["sphere_tests"] or sc.n <= 8          # the grid tests fewer spheres than N per ray


def _mixed_scene(u, seed=4):
    """8000 pebbles, 200 boulders 20 - 40 x their size and the ground: one cell size cannot serve all three."""
    rng = np.random.default_rng(seed)
    r_small, r_big = rng.uniform(0.1, 0.3, 8000), rng.uniform(3.0, 6.0, 200)
    geom = [[0, -1000, 0, 1000]]
    geom += [[x, r, z, r] for x, z, r in zip(rng.uniform(-60, 60, 8000), rng.uniform(-60, 60, 8000), r_small)]
    geom += [[x, r, z, r] for x, z, r in zip(rng.uniform(-60, 60, 200), rng.uniform(-60, 60, 200), r_big)]
    geom = np.array(geom, np.float32)
    n = len(geom)
    mats = rng.integers(0, 3, n).astype(np.uint32); mats[0] = 0
    par = np.concatenate([rng.uniform(0.2, 0.9, (n, 3)), np.where(mats[:, None] == 2, 1.5, rng.uniform(0, 0.3, (n, 1)))], axis=1).astype(np.float32)
    return u.Scene(geom, mats, par)


def test_two_level_grid_mixed_radii(grid, oracle):
    """Scenes whose radii span orders of magnitude get a second, coarse grid level over the spheres too large for the fine one
    (csrc/grid_build.hpp build_grids; the host test checks the level structure).  Same contract: the oracle's hits bit for bit,
    and a render traces exactly the rays the brute-force scan traces."""
    u = grid
    sc = _mixed_scene(u)
    cam = u.Camera.lookat((20, 3, 8), (0, 0.5, 0), vfov=50.0)
    o, d = primary_rays(cam, 480, u.image_height(480))
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    assert (want[2] > 8000).sum() > 2000 and ((want[2] > 0) & (want[2] <= 8000)).sum() > 2000        # boulders and pebbles are both seen
    _check_exact(u.primary_hits(sc, o, d, exact=True), want, "mixed radii grid")
    rng = np.random.default_rng(12)
    o2 = rng.uniform(-65, 65, (50000, 3)).astype(np.float32); o2[:, 1] = np.abs(o2[:, 1]) * 0.15 + 0.05
    d2 = rng.normal(size=(50000, 3)).astype(np.float32)
    _check_exact(u.primary_hits(sc, o2, d2, exact=True), oracle.hits(sc, o2.astype(np.float64), d2.astype(np.float64)), "mixed radii grid, random rays")
    _, _, sg = u.render(sc, cam, 200, 4, 50, seed=6, scan="grid")
    _, _, st = u.render(sc, cam, 200, 4, 50, seed=6, scan="tensor")
    assert sg["scan_mode"] == 3 and st["scan_mode"] == 2 and sg["rays_traced"] == st["rays_traced"]
    assert sg["sphere_tests"] / sg["rays_traced"] < 40          # with the boulders in the always-list it would be > 200
