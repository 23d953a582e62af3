"""Restatement vs the unmodified reference headers, live (only where oracle/_ref was built from
/root/reference).  Larger and more random than the committed goldens; still bit-exact."""
import numpy as np
import pytest


def _rays(rng, n, span):
    o = rng.uniform(-span, span, (n, 3)).astype(np.float32).astype(np.float64)
    d = (rng.normal(size=(n, 3)) * rng.uniform(0.05, 5, (n, 1))).astype(np.float32).astype(np.float64)
    return o, d


@pytest.mark.parametrize("name,span", [("ref4", 3.0), ("final", 14.0), ("dense", 60.0)])
def test_hits_bit_exact(port, ref, u, name, span):
    sc = u.Scene.named(name, seed=3, n=600)
    o, d = _rays(np.random.default_rng(5), 4000, span)
    a, b = port.hits(sc, o, d), ref.hits(sc, o, d)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (a[2] >= 0).sum() > 100


def test_restated_fold_equals_tuple_fold(port, ref, u):
    """The runtime closest-hit loop vs the TRUE tuple-typed hittable_list (hittable_list.hpp:32-58)."""
    o, d = _rays(np.random.default_rng(9), 5000, 2.5)
    t1, n1, i1, f1 = ref.ref4_tuple_hits(o, d)
    for orc in (port, ref):
        t2, n2, i2, f2 = orc.hits(u.Scene.named("ref4"), o, d)
        assert np.array_equal(i1, i2) and np.array_equal(t1, t2) and np.array_equal(n1, n2) and np.array_equal(f1, f2)
    # closed-interval / tie behaviour through the tuple world too
    t, _, i, _ = ref.ref4_tuple_hits([[0, 0, 0]], [[0, 0, -1]], 0.5, 0.5)
    assert i[0] == 0 and t[0] == 0.5


def test_ray_color_matches_tuple_world_up_to_albedo_rounding(port, ref, u):
    s4 = u.Scene.named("ref4")
    for seed in (7, 21, 99):
        a = ref.ref4_tuple_ray_color([0, 0, 0], [0.1 * seed / 99, -0.2, -1], 50, seed)
        b, _ = port.ray_color(s4, [0, 0, 0], [0.1 * seed / 99, -0.2, -1], 50, seed)
        np.testing.assert_allclose(a, b, rtol=5e-7)   # the flattened scene stores float32 albedo
    assert np.array_equal(ref.ref4_tuple_ray_color([0, 0, 0], [0, 0, -1], 50, 7),
                          [0.35704797734894778, 0.18781233703258654, 0.0])


@pytest.mark.parametrize("name,cam", [("ref4", "ref"), ("ref4d", "ref"), ("final", "final")])
def test_render_rows_bit_exact(port, ref, u, name, cam):
    sc, c = u.Scene.named(name, seed=1), u.Camera.named(cam)
    a, ra = port.render_rows(sc, c, 48, 27, 3, 50, 4242, 5, 20)
    b, rb = ref.render_rows(sc, c, 48, 27, 3, 50, 4242, 5, 20)
    assert ra == rb and np.array_equal(a, b)
    assert np.array_equal(port.to_color3b(a, 3), ref.to_color3b(b, 3))


def test_scatter_bit_exact(port, ref):
    rng = np.random.default_rng(1)
    for mt, mp in ((0, (.7, .3, .3, 0)), (1, (.8, .8, .8, 0)), (1, (.8, .6, .2, .45)), (2, (1, 1, 1, 1.5)), (2, (1, 1, 1, 2.4))):
        for k in range(40):
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            din = rng.normal(size=3)
            if np.dot(din, nrm) > 0:
                din = -din
            un = rng.integers(0, 1 << 24, 4) / float(1 << 24)
            a = port.scatter(mt, mp, din, [0, 0, 0], nrm, k & 1, un)
            b = ref.scatter(mt, mp, din, [0, 0, 0], nrm, k & 1, un)
            assert a[0] == b[0]
            if a[0]:
                assert all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))


def test_camera_ray(port, ref, u):
    for cam in ("ref", "final", "dense"):
        c = u.Camera.named(cam)
        for s, t, rx, ry in ((0.5, 0.5, 0, 0), (0.1, 0.9, 0.3, -0.2), (1, 0, -0.7, 0.7)):
            a, b = port.camera_ray(c, s, t, rx, ry), ref.camera_ray(c, s, t, rx, ry)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # the flattened default camera == the reference's camera<T>{} up to float32 rounding of 16/9
    o, d = ref.ref_default_camera_ray(0.25, 0.5)
    o2, d2 = ref.camera_ray(u.Camera.named("ref"), 0.25, 0.5)
    np.testing.assert_allclose(d, d2, rtol=1e-7)
    assert np.array_equal(o, o2)
