"""The CPU restatement (oracle/oracle_port.cpp) against the committed golden vectors, which were
produced by the UNMODIFIED reference headers (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from conftest import golden, unhex


def test_mt19937_uniform_sqrt(port):
    g = golden("kat_primitives.json")
    for k in g["mt19937"]:
        assert port.mt19937(k["seed"], k["n"]) == k["value"]
    for k in g["uniform"]:
        assert port.uniform(k["seed"], k["a"], k["b"], k["n"]) == float.fromhex(k["value"])
    for k in g["sqrt"]:
        assert port.sqrt(float.fromhex(k["x"])) == float.fromhex(k["value"])
    # SURVEY App. B
    assert port.mt19937(42) == 1608637542
    assert port.uniform(42) == 0.79654298428784598


def test_appendix_b_vectors(port, u):
    g = golden("appendix_b_tuple_world.json")
    o, d = unhex(g["o"], (-1, 3)), unhex(g["d"], (-1, 3))
    t, n, i, f = port.hits(u.Scene.named("ref4"), o, d)
    assert i.tolist() == g["id"] == [0, 2, 3, 1, -1, 2]
    assert f.tolist() == g["ff"]
    assert np.array_equal(t, unhex(g["t"]))
    assert np.array_equal(n, unhex(g["n"], (-1, 3)))
    assert t[1] == 0.68665834902296119 and t[3] == 0.62587481340646434 and t[5] == 0.51685579548602179


def test_hit_edge_cases(port, u):
    for c in golden("hit_edge_cases.json")["cases"]:
        g = np.array(c["spheres"], np.float32)
        sc = u.Scene(g, np.zeros(len(g), np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (len(g), 1)))
        tmax = float("inf") if c["tmax"] == "inf" else float.fromhex(c["tmax"])
        t, n, i, f = port.hits(sc, [c["o"]], [c["d"]], float.fromhex(c["tmin"]), tmax)
        assert int(i[0]) == c["id"], c["name"]
        assert int(f[0]) == c["ff"], c["name"]
        assert np.array_equal(t, unhex(c["t"])), c["name"]
        assert np.array_equal(n.ravel(), unhex(c["n"])), c["name"]
    names = {c["name"]: c for c in golden("hit_edge_cases.json")["cases"]}
    assert names["three coincident spheres: later index wins"]["id"] == 2
    assert names["mirror-symmetric equal t: later index wins"]["id"] == 1
    assert names["tangent ray hits"]["id"] == 0


@pytest.mark.parametrize("name", ["ref4", "final"])
def test_random_ray_hits(port, u, name):
    g = golden("random_ray_hits.json")[name]
    sc = u.Scene.named(name, seed=1)
    assert sc.n == g["scene_n"]
    o = np.array(g["o"], np.float32).reshape(-1, 3).astype(np.float64)
    d = np.array(g["d"], np.float32).reshape(-1, 3).astype(np.float64)
    t, n, i, f = port.hits(sc, o, d)
    assert i.tolist() == g["id"] and f.tolist() == g["ff"]
    assert np.array_equal(t, unhex(g["t"])) and np.array_equal(n, unhex(g["n"], (-1, 3)))


def test_scatter_replay(port):
    for c in golden("scatter_replay.json")["cases"]:
        ok, att, oo, do = port.scatter(c["mat_type"], c["mat_param"], np.array(c["d_in"], np.float32), np.array(c["p"], np.float32),
                                       np.array(c["normal"], np.float32), c["ff"], c["uniforms"])
        assert ok == c["scattered"]
        if ok:
            assert np.array_equal(att, unhex(c["att"])) and np.array_equal(oo, unhex(c["o_out"]))
            assert np.array_equal(do, unhex(c["d_out"]))


def test_ray_color_and_render(port, u):
    g = golden("ray_color_and_render.json")
    s4, cam = u.Scene.named("ref4"), u.Camera.named("ref")
    for k in g["ray_color"]:
        rgb, rays = port.ray_color(s4, unhex(k["o"]), unhex(k["d"]), 50, k["seed"])
        assert np.array_equal(rgb, unhex(k["rgb"])) and rays == k["rays"]
    r = g["render"]
    img, rays = port.render_rows(s4, cam, r["w"], r["h"], r["spp"], r["depth"], r["seed"], 0, r["h"])
    assert rays == r["rays"]
    assert np.array_equal(img.ravel(), unhex(r["sum"]))
    assert port.to_color3b(img, r["spp"]).ravel().tolist() == r["rgb8"]
    # row sharding is exact: rows [0,7) + [7,18) == the whole frame
    a, _ = port.render_rows(s4, cam, r["w"], r["h"], r["spp"], r["depth"], r["seed"], 0, 7)
    b, _ = port.render_rows(s4, cam, r["w"], r["h"], r["spp"], r["depth"], r["seed"], 7, r["h"])
    assert np.array_equal(np.concatenate([a, b]), img)
    # threaded whole-frame helper == serial
    c, rc = port.render(s4, cam, r["w"], r["h"], r["spp"], r["depth"], r["seed"], threads=3)
    assert np.array_equal(c, img) and rc == rays


def test_to_color3b_semantics(port):
    # source.cpp:73-83: floor(256*clamp(sqrt(sum/spp), 0, .999)); zero stays zero
    s = np.array([[0.0, 4.0, 1e9], [0.25 * 4, 0.5 * 4, 4 * 0.998001]])
    out = port.to_color3b(s, 4)
    assert out.tolist() == [[0, 255, 255], [128, 181, 255]]
