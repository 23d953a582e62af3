"""-m gpu: the tensor-core candidate scan (tcgen05 TF32 x3 + TMEM, csrc/tc_scan.cuh) against the oracle
and against the FP32 scan.  Both scans only PRE-FILTER spheres (a superset of the reference's
"discriminant >= 0", sphere.hpp:32); the resolve phase is shared, so every closest hit must be the
same sphere: indices bit-exact against the oracle in the FP64-resolved mode, and the two scans
bit-identical to each other in every mode."""
import numpy as np
import pytest

from conftest import golden, primary_rays, unhex
from test_gpu_hits import T_REL_EXACT, _check_exact

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tensor(u):
    u.set_scan_mode("tensor")
    yield u
    u.set_scan_mode("auto")


def _same(a, b, what):
    for x, y, nm in zip(a, b, ("t", "n", "id", "ff")):
        assert np.array_equal(x, y), f"{what}: {nm} differs between the tensor and the FP32 scan"


@pytest.mark.parametrize("name,cam,W", [("ref4", "ref", 400), ("ref4d", "ref", 400), ("final", "final", 1200)])
def test_primary_rays_exact(tensor, oracle, name, cam, W):
    u = tensor
    sc, c = u.Scene.named(name, seed=1), u.Camera.named(cam)
    o, d = primary_rays(c, W, u.image_height(W))
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    got = u.primary_hits(sc, o, d, exact=True)
    _check_exact(got, want, name + " tensor")
    fast = u.primary_hits(sc, o, d, exact=False)
    u.set_scan_mode("fp32")
    _same(got, u.primary_hits(sc, o, d, exact=True), name + " exact")
    _same(fast, u.primary_hits(sc, o, d, exact=False), name + " fast")


@pytest.mark.parametrize("name,span", [("ref4", 3.0), ("final", 14.0)])
def test_random_and_golden_rays_exact(tensor, oracle, name, span):
    u = tensor
    sc = u.Scene.named(name, seed=1)
    g = golden("random_ray_hits.json")[name]
    o = np.array(g["o"], np.float32).reshape(-1, 3); d = np.array(g["d"], np.float32).reshape(-1, 3)
    want = (unhex(g["t"]), unhex(g["n"], (-1, 3)), np.array(g["id"], np.int32), np.array(g["ff"], np.uint8))
    _check_exact(u.primary_hits(sc, o, d, exact=True), want, name + " golden")
    rng = np.random.default_rng(12)
    o = rng.uniform(-span, span, (200000, 3)).astype(np.float32)
    d = (rng.normal(size=(200000, 3)) * rng.uniform(0.05, 5, (200000, 1))).astype(np.float32)
    _check_exact(u.primary_hits(sc, o, d, exact=True), oracle.hits(sc, o.astype(np.float64), d.astype(np.float64)), name)


def test_secondary_rays_from_surfaces_exact(tensor, oracle):
    """Bounce-like rays with origins ON sphere surfaces (incl. the r=1000 ground): the discriminant of the
    sphere a ray leaves is ~0, exactly where a pre-filter must not lose the candidate."""
    u = tensor
    sc, c = u.Scene.named("final", seed=1), u.Camera.named("final")
    o, d = primary_rays(c, 400, 225)
    t, n, i, f = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    hit = i >= 0
    p = (o[hit].astype(np.float64) + d[hit].astype(np.float64) * t[hit][:, None]).astype(np.float32)
    rng = np.random.default_rng(6)
    v = rng.normal(size=p.shape); v /= np.linalg.norm(v, axis=1, keepdims=True)
    d2 = (n[hit] + v).astype(np.float32)
    want = oracle.hits(sc, p.astype(np.float64), d2.astype(np.float64))
    _check_exact(u.primary_hits(sc, p, d2, exact=True), want, "secondary tensor")
    # tangent rays: direction perpendicular to the normal, grazing the sphere they start on
    tang = np.cross(n[hit], v); tang /= np.linalg.norm(tang, axis=1, keepdims=True)
    d3 = tang.astype(np.float32)
    _check_exact(u.primary_hits(sc, p, d3, exact=True), oracle.hits(sc, p.astype(np.float64), d3.astype(np.float64)), "tangent tensor")


@pytest.mark.parametrize("shift", [(300.0, 0.0, 300.0), (3000.0, 50.0, -2000.0)])
def test_far_from_origin_stays_exact(tensor, oracle, shift):
    """The error bound of the TF32-split product grows with |o|^2 and |c|^2; far from the origin the
    bitmaps fill up with false candidates, and the resolve must still return the reference's sphere."""
    u = tensor
    base, c = u.Scene.named("final", seed=1), u.Camera.named("final")
    g = base.geom.copy(); g[:, :3] += np.array(shift, np.float32)
    sc = u.Scene(g, base.mat_type, base.mat_param)
    o, d = primary_rays(c, 240, 135)
    o = o + np.array(shift, np.float32)
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    _check_exact(u.primary_hits(sc, o, d, exact=True), want, f"shift {shift} tensor")


@pytest.mark.parametrize("scale", [1e-3, 1.0, 1e3])
def test_scale_invariance_and_mixed_radii(tensor, oracle, scale):
    """The bound of the TF32-split product is RELATIVE (u * sum|a_i b_i|), so the filter must stay conservative when
    the whole world is scaled by 10^-3 or 10^3, and when radii span four orders of magnitude in one scene."""
    u = tensor
    rng = np.random.default_rng(77)
    n = 700
    radii = 10.0 ** rng.uniform(-2, 1.7, n)                     # 0.01 .. 50
    cen = rng.uniform(-40, 40, (n, 3))
    geom = (np.concatenate([cen, radii[:, None]], axis=1) * scale).astype(np.float32)
    sc = u.Scene(geom, np.zeros(n, np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (n, 1)))
    o = (rng.uniform(-60, 60, (100000, 3)) * scale).astype(np.float32)
    d = rng.normal(size=(100000, 3)).astype(np.float32)
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64), 0.001 * scale, float("inf"))
    assert (want[2] >= 0).mean() > 0.3
    _check_exact(u.primary_hits(sc, o, d, 0.001 * scale, float("inf"), exact=True), want, f"scale {scale}")
    # rays leaving the surfaces (discriminant of the own sphere ~ 0 at every scale)
    t, nrm, i, f = want
    hit = i >= 0
    p = (o[hit].astype(np.float64) + d[hit].astype(np.float64) * t[hit][:, None]).astype(np.float32)
    v = rng.normal(size=p.shape); v /= np.linalg.norm(v, axis=1, keepdims=True)
    d2 = (nrm[hit] + v).astype(np.float32)
    want2 = oracle.hits(sc, p.astype(np.float64), d2.astype(np.float64), 0.001 * scale, float("inf"))
    _check_exact(u.primary_hits(sc, p, d2, 0.001 * scale, float("inf"), exact=True), want2, f"scale {scale} secondary")


def test_edge_cases_golden(tensor):
    u = tensor
    for c in golden("hit_edge_cases.json")["cases"]:
        g = np.array(c["spheres"], np.float32)
        sc = u.Scene(g, np.zeros(len(g), np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (len(g), 1)))
        tmax = float("inf") if c["tmax"] == "inf" else float.fromhex(c["tmax"])
        t, n, i, f = u.primary_hits(sc, [c["o"]], [c["d"]], float.fromhex(c["tmin"]), tmax, exact=True)
        assert int(i[0]) == c["id"], c["name"]
        if c["id"] >= 0:
            assert int(f[0]) == c["ff"], c["name"]
            assert abs(t[0] - unhex(c["t"])[0]) <= T_REL_EXACT * abs(unhex(c["t"])[0]), c["name"]


@pytest.mark.parametrize("n", [0, 1, 2, 127, 128, 129, 500, 1023, 1024, 1025, 1151, 1152, 1153, 1536, 1537, 2049, 3000])
def test_scene_sizes_around_the_chunk_and_row_limits(tensor, oracle, n):
    """Sphere counts around the 128-column MMA chunk, the residency limit (1152 rows) and, above it, the 512-row
    slices of the swept operand (pad rows never hit)."""
    u = tensor
    rng = np.random.default_rng(100 + n)
    geom = np.concatenate([rng.uniform(-6, 6, (n, 3)), rng.uniform(0.1, 0.6, (n, 1))], axis=1).astype(np.float32)
    sc = u.Scene(geom, np.zeros(n, np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (n, 1)))
    o = rng.uniform(-8, 8, (20000, 3)).astype(np.float32)
    d = rng.normal(size=(20000, 3)).astype(np.float32)
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    _check_exact(u.primary_hits(sc, o, d, exact=True), want, f"n={n}")
    if n > 8:
        assert (want[2] >= 0).any() and want[2].max() >= n - 1 - n // 4      # the last spheres (next to the pad rows) do get hit


def test_tie_rule_later_index_wins(tensor):
    """Coincident spheres in different 32-sphere bitmap words and MMA chunks: the LAST one wins (hittable_list.hpp:32-58)."""
    u = tensor
    n = 300
    geom = np.tile(np.array([[50.0, 50.0, 50.0, 0.1]], np.float32), (n, 1))
    for k in (3, 40, 131, 299):
        geom[k] = (0, 0, -1, 0.5)
    sc = u.Scene(geom, np.zeros(n, np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (n, 1)))
    for exact in (True, False):
        t, nn, i, f = u.primary_hits(sc, [[0, 0, 0]], [[0, 0, -1]], exact=exact)
        assert int(i[0]) == 299 and abs(t[0] - 0.5) < 1e-6


def test_non_finite_geometry_is_rejected(u):
    cam = u.Camera.named("ref")
    for bad in (np.nan, np.inf, 1e30):
        geom = np.array([[0, 0, -1, 0.5], [bad, 0, -1, 0.5]], np.float32)
        sc = u.Scene(geom, np.zeros(2, np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (2, 1)))
        with pytest.raises(u.UecrtError):
            u.render(sc, cam, 32, 1, 5, seed=1)
        with pytest.raises(u.UecrtError):
            u.primary_hits(sc, [[0, 0, 0]], [[0, 0, -1]])


def _random_scene(u, n, seed, span=6.0):
    rng = np.random.default_rng(seed)
    geom = np.concatenate([rng.uniform(-span, span, (n, 3)), rng.uniform(0.1, 0.4, (n, 1))], axis=1).astype(np.float32)
    mats = rng.integers(0, 3, n).astype(np.uint32)
    par = np.concatenate([rng.uniform(0.2, 0.9, (n, 3)), np.where(mats[:, None] == 2, 1.5, rng.uniform(0, 0.3, (n, 1)))], axis=1).astype(np.float32)
    return u.Scene(geom, mats, par)


def test_limits_reported(u):
    r = u.tensor_scan_resident_spheres()          # operand + two quad-flag bitmaps + the shade side's tables per sphere row
    assert r % 128 == 0 and 1024 <= r <= 1536
    assert u.tensor_scan_max_spheres() == 1 << 22       # the scene size limit; the FP32 scan alone stops at 65 534 (16-bit candidate lists)


@pytest.mark.parametrize("n", [1025, 2048, 5000])
@pytest.mark.parametrize("exact", [False, True])
def test_swept_operand_hits_identical_to_fp32_scan(u, n, exact):
    """Scenes above the residency limit: the sphere operand is swept in 512-row slices, a ray's candidates
    arrive pass by pass and fold into one closest hit.  t, normal, id, front_face must equal the FP32 scan's
    bit for bit, including ties between coincident spheres that sit in different slices (later index wins)."""
    sc = _random_scene(u, n, 40 + n, span=9.0)
    g = sc.geom.copy()
    for k in (7, 600, 1024, n - 1):                       # coincident spheres across slice boundaries
        g[k] = (0.5, -0.25, -3.0, 0.5)
    sc = u.Scene(g, sc.mat_type, sc.mat_param)
    rng = np.random.default_rng(n)
    o = rng.uniform(-10, 10, (60000, 3)).astype(np.float32)
    d = (rng.normal(size=(60000, 3)) * rng.uniform(0.05, 5, (60000, 1))).astype(np.float32)
    o[:100] = 0; d[:100] = (0.5, -0.25, -3.0) + rng.normal(scale=0.02, size=(100, 3))
    out = {}
    for mode in ("fp32", "tensor"):
        u.set_scan_mode(mode)
        try:
            out[mode] = u.primary_hits(sc, o, d, exact=exact)
        finally:
            u.set_scan_mode("auto")
    a, b = out["fp32"], out["tensor"]
    assert (a[2] >= 0).sum() > 1000 and (a[2] == n - 1).sum() > 0 and not np.isin(a[2], [k for k in (7, 600, 1024) if k != n - 1]).any()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_swept_operand_at_the_scene_size_limit(u):
    """65 534 spheres (the FP32 scan's limit, so the largest scene on which the two scans can be compared): 128 slices of
    512 rows, an 8 MB operand streamed from L2."""
    n = 65534
    sc = _random_scene(u, n, 77, span=40.0)
    rng = np.random.default_rng(78)
    o = rng.uniform(-45, 45, (3000, 3)).astype(np.float32)
    d = rng.normal(size=(3000, 3)).astype(np.float32)
    out = {}
    for mode in ("fp32", "tensor"):
        u.set_scan_mode(mode)
        try:
            out[mode] = u.primary_hits(sc, o, d, exact=True)
        finally:
            u.set_scan_mode("auto")
    assert (out["fp32"][2] >= 0).sum() > 500 and out["fp32"][2].max() > n - n // 8
    for x, y in zip(out["fp32"], out["tensor"]):
        assert np.array_equal(x, y)


def test_swept_operand_render_takes_the_same_paths(u, oracle):
    n = 3000
    sc, cam = u.Scene.named("dense", seed=1, n=n), u.Camera.named("dense")
    ra, aa, sa = u.render(sc, cam, 320, 8, 50, seed=6, return_accum=True)                  # automatic: swept tensor scan
    rb, ab, sb = u.render(sc, cam, 320, 8, 50, seed=6, return_accum=True, scan="fp32")
    assert sa["scan_mode"] == 2 and sb["scan_mode"] == 1 and sa["k1_threads"] == 768
    assert sa["rays_traced"] == sb["rays_traced"] and sa["paths"] == sb["paths"] == 320 * u.image_height(320) * 8
    assert np.all(aa[..., 3] == 8) and np.abs(aa[..., :3] - ab[..., :3]).max() <= 1e-3
    for exact in (True,):
        _, ae, se = u.render(sc, cam, 160, 4, 50, seed=6, return_accum=True, exact=exact, scan="tensor")
        _, af, sf = u.render(sc, cam, 160, 4, 50, seed=6, return_accum=True, exact=exact, scan="fp32")
        assert se["rays_traced"] == sf["rays_traced"] and np.abs(ae - af).max() <= 1e-3
    # primary hits of the swept scan against the oracle (exact mode: index and front face bit-exact)
    rng = np.random.default_rng(3)
    o = np.concatenate([rng.uniform(-50, 50, (20000, 1)), rng.uniform(0.05, 6, (20000, 1)), rng.uniform(-50, 50, (20000, 1))], axis=1).astype(np.float32)
    d = rng.normal(size=(20000, 3)).astype(np.float32)
    u.set_scan_mode("tensor")
    try:
        got = u.primary_hits(sc, o, d, exact=True)
    finally:
        u.set_scan_mode("auto")
    _check_exact(got, oracle.hits(sc, o.astype(np.float64), d.astype(np.float64)), "dense 3000 swept")


def test_forced_tiles_still_select_the_fp32_scan_on_large_scenes(u):
    sc, cam = _random_scene(u, 1500, 8), u.Camera.named("ref")
    u.set_tile_spheres(256)
    try:
        _, _, st = u.render(sc, cam, 64, 4, 5, seed=1)
        assert st["scan_mode"] == 1 and st["tiles"] > 1
    finally:
        u.set_tile_spheres(0)


@pytest.mark.parametrize("name,cam,W,spp", [("ref4", "ref", 400, 32), ("ref4d", "ref", 400, 32), ("final", "final", 400, 32)])
@pytest.mark.parametrize("exact", [False, True])
def test_render_identical_to_fp32_scan(u, name, cam, W, spp, exact):
    """Same Philox streams, same candidates after the resolve: every path takes the same bounces.  The two
    kernels add the finished paths into the accumulation buffer in a different order, so the float
    sums may differ in the last bits; ray counts, path counts and sample counts must be equal."""
    sc, c = u.Scene.named(name, seed=1), u.Camera.named(cam)
    ra, aa, sa = u.render(sc, c, W, spp, 50, seed=9, exact=exact, return_accum=True, scan="tensor")
    rb, ab, sb = u.render(sc, c, W, spp, 50, seed=9, exact=exact, return_accum=True, scan="fp32")
    assert sa["scan_mode"] == 2 and sb["scan_mode"] == 1
    assert sa["rays_traced"] == sb["rays_traced"] and sa["paths"] == sb["paths"] == W * u.image_height(W) * spp
    assert np.array_equal(aa[..., 3], ab[..., 3]) and np.all(aa[..., 3] == spp)
    assert np.abs(aa[..., :3] - ab[..., :3]).max() <= 1e-4 * spp
    assert (ra != rb).mean() < 1e-4                                        # 8-bit pixels on a rounding edge


def test_auto_mode_uses_tensor_scan_and_stats(u):
    sc, c = u.Scene.named("final", seed=1), u.Camera.named("final")
    _, _, st = u.render(sc, c, 200, 8, 50, seed=2)
    # resident operand: the self-scan pipeline (csrc/tc_self.cuh), 6 groups of 128 paths + the issuer warpgroup
    assert st["scan_mode"] == 2 and st["k1_threads"] == 896 and st["k1_blocks"] == st["sm_count"]
    assert st["rays_traced"] > st["paths"] == 200 * u.image_height(200) * 8
    _, _, st4 = u.render(u.Scene.named("ref4", seed=1), u.Camera.named("ref"), 200, 8, 50, seed=2)
    assert st4["scan_mode"] == 1           # below 96 spheres the FP32 scan is faster and the automatic mode uses it
    u.set_tile_spheres(64)                 # forced shared-memory tiles are a request for the FP32 scan's ring
    try:
        _, _, st = u.render(sc, c, 200, 8, 50, seed=2)
        assert st["scan_mode"] == 1 and st["tiles"] > 1
    finally:
        u.set_tile_spheres(0)


@pytest.mark.parametrize("W", [2, 7, 33, 250])
def test_ragged_edges_and_tiny_frames(u, W):
    """Frames smaller than one shade group and ragged 8x4 tile edges: dead rows of an MMA tile carry P = +1e30."""
    sc, c = u.Scene.named("ref4d", seed=1), u.Camera.named("ref")
    ra, aa, sa = u.render(sc, c, W, 16, 50, seed=4, return_accum=True, scan="tensor")
    rb, ab, sb = u.render(sc, c, W, 16, 50, seed=4, return_accum=True, scan="fp32")
    assert sa["rays_traced"] == sb["rays_traced"] and sa["paths"] == sb["paths"]
    assert np.all(aa[..., 3] == 16) and np.abs(aa[..., :3] - ab[..., :3]).max() <= 2e-3


def test_depth_and_sample_shards(u):
    sc, c = u.Scene.named("final", seed=1), u.Camera.named("final")
    for depth in (0, 1, 2):
        _, aa, sa = u.render(sc, c, 120, 8, depth, seed=3, return_accum=True, scan="tensor")
        _, ab, sb = u.render(sc, c, 120, 8, depth, seed=3, return_accum=True, scan="fp32")
        assert sa["rays_traced"] == sb["rays_traced"] and np.abs(aa - ab).max() <= 1e-3
        if depth == 0:
            assert sa["rays_traced"] == 0 and np.all(aa[..., :3] == 0)
    _, whole, _ = u.render(sc, c, 160, 24, 50, seed=5, return_accum=True, scan="tensor")
    parts = [u.render(sc, c, 160, 24, 50, seed=5, return_accum=True, scan="tensor", sample_begin=b, sample_count=8)[1] for b in (0, 8, 16)]
    assert np.abs(sum(parts) - whole).max() <= 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("name,cam,n", [("final", "final", None), ("dense", "dense", 130), ("dense", "dense", 1100)])
def test_both_tensor_pipelines_trace_the_same_paths(u, name, cam, n, exact):
    """The self-scan pipeline (every shade warp reads its own accumulators, csrc/tc_self.cuh) and the scan-warp pipeline
    (csrc/tc_scan.cuh) are two schedules of the same candidate scan: same flags, same resolve, same Philox streams, so
    every path takes the same bounces; only the order of the atomic adds into the accumulation buffer differs."""
    sc = u.Scene.named(name, seed=1) if n is None else u.Scene.named(name, seed=1, n=n)
    c = u.Camera.named(cam)
    out = {}
    try:
        for p in ("self", "scanwarps"):
            u.set_tensor_pipeline(p)
            out[p] = u.render(sc, c, 240, 12, 50, seed=5, exact=exact, return_accum=True, scan="tensor")
    finally:
        u.set_tensor_pipeline("self")
    (ra, aa, sa), (rb, ab, sb) = out["self"], out["scanwarps"]
    resident_self = sc.n <= 1152
    assert sa["k1_threads"] == (896 if resident_self else 768) and sb["k1_threads"] == 768
    assert sa["rays_traced"] == sb["rays_traced"] and sa["paths"] == sb["paths"] == 240 * u.image_height(240) * 12
    assert np.array_equal(aa[..., 3], ab[..., 3]) and np.all(aa[..., 3] == 12)
    assert np.abs(aa[..., :3] - ab[..., :3]).max() <= 1e-4 * 12
    assert (ra != rb).mean() < 1e-4


def test_scene_above_the_16_bit_list_limit(u, oracle):
    """70 000 spheres: more than the FP32 scan's 16-bit candidate lists can index (65 534).  The automatic mode takes the
    swept tensor scan (candidate bitmaps) and the grid holds 32-bit items: both still return the reference's hits."""
    n = 70000
    sc, cam = u.Scene.named("dense", seed=3, n=n), u.Camera.named("dense")
    assert sc.n == n and u.tensor_scan_max_spheres() >= n
    rng = np.random.default_rng(17)
    o = rng.uniform(-60, 60, (1500, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1]) * 0.2 + 0.3
    d = rng.normal(size=(1500, 3)).astype(np.float32)
    want = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    assert (want[2] >= 0).sum() > 300 and want[2].max() > 65535
    _check_exact(u.primary_hits(sc, o, d), want, "auto (tensor, swept)")
    u.set_scan_mode("grid")
    try:
        _check_exact(u.primary_hits(sc, o, d), want, "grid")
    finally:
        u.set_scan_mode("auto")
    u.set_scan_mode("fp32")
    try:
        with pytest.raises(u.UecrtError):
            u.primary_hits(sc, o, d)                       # forced FP32 scan: refused, not silently wrong
    finally:
        u.set_scan_mode("auto")
    _, _, sa = u.render(sc, cam, 96, 2, 50, seed=4)
    _, _, sg = u.render(sc, cam, 96, 2, 50, seed=4, scan="grid")
    assert sa["scan_mode"] == 2 and sg["scan_mode"] == 3 and sa["rays_traced"] == sg["rays_traced"] and sa["paths"] == 96 * u.image_height(96) * 2


def test_self_scan_pipeline_stress(u):
    """The self-scan pipeline's barrier protocol (lanes, per-lane termination, a work queue shorter or longer than the
    launch) over many small launches: 1 to 9 chunks, images from a few pixels (most groups never get a path) to ragged
    sizes, 1 to 5 spp, depth 0 to 50.  Every launch must end (the test would hang otherwise), count every sample exactly
    once and trace exactly the rays the FP32 scan traces.  (compute-sanitizer is closed on the pool; this is the race check.)"""
    rng = np.random.default_rng(2026)
    cam = u.Camera.named("dense")
    for k in range(36):
        n = int(rng.choice([96, 127, 128, 129, 255, 300, 512, 513, 777, 1024, 1100, 1152]))
        W = int(rng.choice([8, 16, 33, 64, 100, 177]))
        spp = int(rng.integers(1, 6))
        depth = int(rng.choice([0, 1, 2, 7, 50]))
        sc = u.Scene.named("dense", seed=10 + k, n=n)
        exact = bool(k % 3 == 0)
        _, a, sa = u.render(sc, cam, W, spp, depth, seed=k, exact=exact, return_accum=True, scan="tensor")
        _, b, sb = u.render(sc, cam, W, spp, depth, seed=k, exact=exact, return_accum=True, scan="fp32")
        tag = (n, W, spp, depth, exact)
        assert sa["k1_threads"] == 896 and sa["scan_mode"] == 2, tag
        assert np.all(a[..., 3] == spp) and sa["paths"] == sb["paths"] == W * u.image_height(W) * spp, tag
        assert sa["rays_traced"] == sb["rays_traced"], tag
        assert np.abs(a[..., :3] - b[..., :3]).max() <= 1e-4 * spp, tag
