"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/liboracle_ref.so, built by
`make -C oracle ref` from the headers under /root/reference).  Run in the build container only:
    python tests/golden/make_golden.py
The fixtures are what pins oracle/oracle_port.cpp (and, on the GPU box where /root/reference does
not exist, the CUDA path).  Doubles are stored as hex strings so they round-trip bit-exactly.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import uecraytracing_b200 as u  # noqa: E402  (host scene builders only; no GPU use)
from oracle import pyoracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ref = pyoracle.load("ref")
assert ref.kind == "reference"


def hx(a):
    return [float(x).hex() for x in np.asarray(a, np.float64).ravel()]


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=0, separators=(",", ":"))
    print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")


rng = np.random.default_rng(20261018)

# ---- 1. primitives: mt19937, uniform_real_distribution, math::sqrt (random.hpp, math.hpp)
kat = {
    "mt19937": [{"seed": s, "n": n, "value": ref.mt19937(s, n)} for s in (42, 0, 5489, 4294967295) for n in (0, 1, 623, 624, 1000)],
    "uniform": [{"seed": s, "a": a, "b": b, "n": n, "value": float(ref.uniform(s, a, b, n)).hex()}
                for s in (42, 7) for (a, b) in ((0.0, 1.0), (-1.0, 1.0), (0.01, 0.99)) for n in (0, 5)],
    "sqrt": [{"x": float(x).hex(), "value": float(ref.sqrt(x)).hex()} for x in (2.0, 0.25, 1e-12, 3.0, 19000.0, 1e300)],
}
dump("kat_primitives.json", kat)

# ---- 2. SURVEY App. B: the TRUE tuple-typed hittable_list on the reference camera's rays
uv = [(0.5, 0.5), (0.25, 0.5), (0.75, 0.5), (0.5, 0.1), (0.5, 0.9), (0.02, 0.55)]
o = np.array([ref.ref_default_camera_ray(a, b)[0] for a, b in uv]); d = np.array([ref.ref_default_camera_ray(a, b)[1] for a, b in uv])
t, n, i, f = ref.ref4_tuple_hits(o, d)
appb = {"uv": uv, "o": hx(o), "d": hx(d), "t": hx(t), "n": hx(n), "id": i.tolist(), "ff": f.tolist(),
        "ray_color_seed7_center": hx(ref.ref4_tuple_ray_color([0, 0, 0], [0, 0, -1], 50, 7))}
dump("appendix_b_tuple_world.json", appb)

# ---- 3. adversarial hit cases (SURVEY App. B "other known answers"): ties, closed interval, inside, tangent
def sc(spheres):
    g = np.array([[*c, r] for c, r in spheres], np.float32)
    return u.Scene(g, np.zeros(len(g), np.uint32), np.tile(np.array([[.5, .5, .5, 0]], np.float32), (len(g), 1)))

cases = []
def case(name, spheres, o, d, tmin, tmax):
    t, n, i, f = ref.hits(sc(spheres), [o], [d], tmin, tmax)
    cases.append({"name": name, "spheres": [[*c, r] for c, r in spheres], "o": o, "d": d, "tmin": float(tmin).hex(),
                  "tmax": "inf" if np.isinf(tmax) else float(tmax).hex(), "t": hx(t), "n": hx(n), "id": int(i[0]), "ff": int(f[0])})
inf = float("inf")
case("three coincident spheres: later index wins", [((0, 0, -1), .5)] * 3, [0, 0, 0], [0, 0, -1], 0.001, inf)
case("origin inside", [((0, 0, 0), 2.0)], [0, 0, 0], [0, 0, -1], 0.001, inf)
case("root == t_min accepted", [((0, 0, -1.5), .5)], [0, 0, 0], [0, 0, -1], 1.0, 2.0)
case("t_min just above near root -> far root", [((0, 0, -1.5), .5)], [0, 0, 0], [0, 0, -1], 1.0000001, 2.0)
case("root == t_max accepted", [((0, 0, -1.5), .5)], [0, 0, 0], [0, 0, -1], 0.0, 1.0)
case("t_max below near root -> miss", [((0, 0, -1.5), .5)], [0, 0, 0], [0, 0, -1], 0.0, 0.999)
case("tangent ray hits", [((0, 0, -1), .5)], [0.5, 0, 0], [0, 0, -1], 0.001, inf)
case("negative radius flips outward normal", [((0, 0, -1), -.5)], [0, 0, 0], [0, 0, -1], 0.001, inf)
case("unnormalised direction: t in units of |d|", [((0, 0, -10), 1.0)], [0, 0, 0], [0, 0, -4], 0.001, inf)
case("sphere behind the origin -> miss", [((0, 0, 5), 1.0)], [0, 0, 0], [0, 0, -1], 0.001, inf)
case("two spheres, nearer later in list", [((0, 0, -5), 1.0), ((0, 0, -2), .5)], [0, 0, 0], [0, 0, -1], 0.001, inf)
case("two spheres, nearer first in list", [((0, 0, -2), .5), ((0, 0, -5), 1.0)], [0, 0, 0], [0, 0, -1], 0.001, inf)
case("mirror-symmetric equal t: later index wins", [((-1, 0, -3), 1.5), ((1, 0, -3), 1.5)], [0, 0, 0], [0, 0, -1], 0.001, inf)
dump("hit_edge_cases.json", {"cases": cases})

# ---- 4. random rays against the reference scene and a 64-sphere random scene (restated fold vs headers)
def rand_rays(n, span):
    o = rng.uniform(-span, span, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32) * rng.uniform(0.1, 4.0, (n, 1)).astype(np.float32)
    return o, d

hits = {}
for name, scene, span in (("ref4", u.Scene.named("ref4"), 3.0), ("final", u.Scene.named("final", seed=1), 14.0)):
    o, d = rand_rays(256, span)
    t, n, i, f = ref.hits(scene, o.astype(np.float64), d.astype(np.float64))
    hits[name] = {"o": o.ravel().tolist(), "d": d.ravel().tolist(), "t": hx(t), "n": hx(n), "id": i.tolist(), "ff": f.tolist(),
                  "scene_n": scene.n}
dump("random_ray_hits.json", hits)

# ---- 5. scatter with replayed uniforms: lambertian / metal are the reference's own code
scat = []
for mt, mp in ((0, (.7, .3, .3, 0)), (1, (.8, .8, .8, 0)), (1, (.8, .6, .2, .3)), (2, (1, 1, 1, 1.5))):
    for k in range(6):
        nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
        din = rng.normal(size=3) * rng.uniform(0.2, 3)
        if np.dot(din, nrm) > 0: din = -din      # normal faces against the ray, as set_face_normal guarantees
        p = rng.uniform(-2, 2, 3)
        ff = int(k % 2 == 0) if mt == 2 else 1
        un = (rng.integers(0, 1 << 24, 4) / float(1 << 24)).tolist()
        nrm32, din32, p32 = nrm.astype(np.float32), din.astype(np.float32), p.astype(np.float32)
        ok, att, oo, do = ref.scatter(mt, mp, din32.astype(np.float64), p32.astype(np.float64), nrm32.astype(np.float64), ff, un)
        scat.append({"mat_type": mt, "mat_param": list(mp), "d_in": din32.tolist(), "p": p32.tolist(), "normal": nrm32.tolist(),
                     "ff": ff, "uniforms": un, "scattered": ok, "att": hx(att), "o_out": hx(oo), "d_out": hx(do)})
dump("scatter_replay.json", {"cases": scat})

# ---- 6. ray_color with seeded mt19937 on the flattened reference scene + a tiny whole-frame render
s4 = u.Scene.named("ref4"); cam = u.Camera.named("ref")
rc = []
for seed, (a, b) in ((7, (0.5, 0.5)), (11, (0.3, 0.4)), (13, (0.8, 0.2)), (17, (0.5, 0.95))):
    o_, d_ = ref.camera_ray(cam, a, b)
    rgb, rays = ref.ray_color(s4, o_, d_, 50, seed)
    rc.append({"seed": seed, "s": a, "t": b, "o": hx(o_), "d": hx(d_), "rgb": hx(rgb), "rays": rays})
W, H, SPP = 32, 18, 8
img, rays = ref.render_rows(s4, cam, W, H, SPP, 50, 12345, 0, H)
dump("ray_color_and_render.json", {"ray_color": rc, "render": {"w": W, "h": H, "spp": SPP, "depth": 50, "seed": 12345,
                                                               "sum": hx(img), "rays": rays, "rgb8": ref.to_color3b(img, SPP).ravel().tolist()}})

# ---- 7. converged low-res oracle images (the GPU box has no /root/reference): linear means per pixel
for name, scene, camn, W, SPP in (("ref4", s4, "ref", 64, 1024), ("final", u.Scene.named("final", seed=1), "final", 64, 1024)):
    cam_ = u.Camera.named(camn); H = u.image_height(W)
    a, ra = ref.render(scene, cam_, W, H, SPP, 50, 1)
    b, rb = ref.render(scene, cam_, W, H, SPP, 50, 777777)
    mean = ((a + b) / (2 * SPP)).astype(np.float32)
    mse_ab = float(np.mean(((a - b) / SPP) ** 2))
    dump(f"converged_{name}_{W}.json", {"w": W, "h": H, "spp_total": 2 * SPP, "depth": 50, "scene": name, "camera": camn,
                                        "mean_rgb_f32": mean.ravel().tolist(), "mse_between_halves": mse_ab,
                                        "rays_per_sample": (ra + rb) / (2.0 * W * H * SPP)})
