"""The tensor-core candidate scan must flag a SUPERSET of the spheres the reference's `discriminant >= 0` accepts
(sphere.hpp:27-32): a sphere it drops is a silently wrong hit id.  The scan's verdict is read back before any resolve
(uecrt_debug_tensor_candidates) on rays built to sit ON the decision boundary: exactly tangent and near-tangent to
every sphere, at scene scales 1e-3 / 1 / 1e3, 3000 units from the origin, and on the final scene with its r = 1000
ground sphere.  Truth is the reference's discriminant evaluated in double, in its operation order, on the same float
inputs the device gets."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def reference_discriminant(o, d, geom):
    """sphere.hpp:27-31 for every (ray, sphere) pair, IEEE double, no contraction (numpy evaluates op by op)."""
    o = o.astype(np.float64); d = d.astype(np.float64); g = geom.astype(np.float64)
    ocx = o[:, None, 0] - g[None, :, 0]; ocy = o[:, None, 1] - g[None, :, 1]; ocz = o[:, None, 2] - g[None, :, 2]
    dx, dy, dz = d[:, None, 0], d[:, None, 1], d[:, None, 2]
    a = (dx * dx + dy * dy) + dz * dz
    half_b = (ocx * dx + ocy * dy) + ocz * dz
    c = ((ocx * ocx + ocy * ocy) + ocz * ocz) - g[None, :, 3] * g[None, :, 3]
    return half_b * half_b - a * c, a * ((ocx * ocx + ocy * ocy) + ocz * ocz)


def tangent_rays(geom, per_sphere, rng, spread, dlen=(0.05, 5.0)):
    """per_sphere rays for every sphere whose line touches it: exactly tangent in real arithmetic, then perturbed by
    0, +-1e-7, +-1e-6 of the tangent angle's sine, then rounded to float (so the double discriminant on the float
    inputs scatters around zero at the rounding level)."""
    n = len(geom)
    c = np.repeat(geom[:, :3].astype(np.float64), per_sphere, axis=0)
    r = np.repeat(np.abs(geom[:, 3]).astype(np.float64), per_sphere)
    m = len(c)
    dirs = rng.normal(size=(m, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dist = r * (1.0 + 10.0 ** rng.uniform(-2, np.log10(spread), m))          # from just above the surface to far away
    o = c + dirs * dist[:, None]
    o = o.astype(np.float32).astype(np.float64)                               # the origin the device sees
    w = c - o; L = np.linalg.norm(w, axis=1); wh = w / L[:, None]
    t = rng.normal(size=(m, 3)); t -= (t * wh).sum(1, keepdims=True) * wh; t /= np.linalg.norm(t, axis=1, keepdims=True)
    eps = rng.choice([0.0, 0.0, 1e-7, -1e-7, 1e-6, -1e-6, 3e-8, -3e-8], m)
    s = np.clip(r / L * (1.0 + eps), 0.0, 1.0)
    d = np.sqrt(1.0 - s * s)[:, None] * wh + s[:, None] * t
    d *= rng.uniform(dlen[0], dlen[1], (m, 1))                                # the reference never normalises the direction
    o32, d32 = o.astype(np.float32), d.astype(np.float32)
    # walk the direction's largest component a few float ulps either way and keep, for every second ray, the step whose
    # double discriminant against the ray's own sphere is the smallest non-negative one: as close to the boundary as
    # float inputs get, on the accepting side
    g = np.repeat(geom, per_sphere, axis=0)
    k = np.argmax(np.abs(d32), axis=1)
    best, best_disc = d32.copy(), np.full(m, np.inf)
    for step in range(-8, 9):
        cand = d32.copy()
        v = cand[np.arange(m), k]
        for _ in range(abs(step)):
            v = np.nextafter(v, np.float32(np.inf) if step > 0 else np.float32(-np.inf))
        cand[np.arange(m), k] = v
        disc = _disc_own(o32, cand, g)
        better = (disc >= 0) & (disc < best_disc)
        best[better] = cand[better]; best_disc[better] = disc[better]
    pick = (np.arange(m) % 2 == 0) & np.isfinite(best_disc)
    d32[pick] = best[pick]
    return o32, d32


def _disc_own(o, d, g):
    o = o.astype(np.float64); d = d.astype(np.float64); g = g.astype(np.float64)
    oc = o - g[:, :3]
    a = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    hb = (oc[:, 0] * d[:, 0] + oc[:, 1] * d[:, 1]) + oc[:, 2] * d[:, 2]
    c = ((oc[:, 0] * oc[:, 0] + oc[:, 1] * oc[:, 1]) + oc[:, 2] * oc[:, 2]) - g[:, 3] * g[:, 3]
    return hb * hb - a * c


def random_scene(u, n, rng, scale, offset):
    geom = np.concatenate([rng.uniform(-6, 6, (n, 3)), rng.uniform(0.1, 0.6, (n, 1))], axis=1) * scale
    geom[:, :3] += offset
    geom = geom.astype(np.float32)
    return u.Scene(geom, np.zeros(n, np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (n, 1)))


def check_superset(u, sc, o, d, label):
    flags, per = u.tensor_candidates(sc, o, d)
    assert per in (1, 4)
    disc, scale2 = reference_discriminant(o, d, sc.geom)
    hit = disc >= 0
    dropped = hit & (flags == 0)
    near = hit & (disc <= 2.0 ** -24 * scale2)            # on the boundary: below ONE unit u |oc|^2 |d|^2 of the scan's error bound
    assert not dropped.any(), (f"{label}: {int(dropped.sum())} of {int(hit.sum())} reference candidates were dropped by the tensor scan; "
                               f"first (ray, sphere) {np.argwhere(dropped)[0].tolist()}, disc {disc[dropped][0]:.3e}")
    return int(hit.sum()), int(near.sum()), int(flags.sum()), flags.size


@pytest.mark.parametrize("scale,offset", [(1e-3, 0.0), (1.0, 0.0), (1e3, 0.0), (1.0, 3000.0), (1e-2, 30.0)])
def test_tangent_rays_are_always_candidates(u, scale, offset):
    rng = np.random.default_rng(int(scale * 1000) + int(offset))
    n, per = 160, 110                                      # 160 spheres (two MMA chunks, 40 quads) x 110 tangent rays x 160 spheres = 2.8 M pairs
    sc = random_scene(u, n, rng, scale, offset)
    o, d = tangent_rays(sc.geom, per, rng, spread=50.0)
    hits, near, flagged, pairs = check_superset(u, sc, o, d, f"scale {scale} offset {offset}")
    assert pairs >= 2_500_000 and hits > 10_000
    assert near > 2_000, near                              # thousands of pairs sit on the boundary itself
    if offset < 100 * scale:                               # near the origin it is still a filter, not "everything is a candidate";
        assert flagged < 0.5 * pairs                       # far out the margins (~ u |c|^2) exceed r^2 and it passes everything on


def test_tangent_rays_on_the_final_scene(u):
    """Ground sphere r = 1000 at y = -1000 next to r = 0.2 spheres: the lifted vectors span 1e6 : 1e-2 in one product."""
    rng = np.random.default_rng(5)
    sc = u.Scene.named("final", seed=1)
    o, d = tangent_rays(sc.geom, 12, rng, spread=30.0)
    o2, d2 = tangent_rays(sc.geom[:1], 4000, rng, spread=0.05)          # grazing the ground from points a few units above it
    hits, near, flagged, pairs = check_superset(u, sc, np.concatenate([o, o2]), np.concatenate([d, d2]), "final scene")
    assert pairs > 4_000_000 and near > 1_000


def test_exact_hits_on_tangent_rays_equal_the_oracle(u, oracle):
    """End to end on the same adversarial rays: the FP64-resolved closest hit (tensor scan, quad re-test, exact resolve)
    is the reference's sphere for every ray."""
    rng = np.random.default_rng(9)
    sc = random_scene(u, 300, rng, 1.0, 0.0)
    o, d = tangent_rays(sc.geom, 40, rng, spread=20.0)
    u.set_scan_mode("tensor")
    try:
        t, nrm, idx, ff = u.primary_hits(sc, o, d, exact=True)
    finally:
        u.set_scan_mode("auto")
    to, no, io, fo = oracle.hits(sc, o.astype(np.float64), d.astype(np.float64))
    assert np.array_equal(idx, io) and np.array_equal(ff, fo)
    h = io >= 0
    assert h.sum() > 5000 and np.max(np.abs(t[h] - to[h]) / np.abs(to[h])) <= 1e-9
