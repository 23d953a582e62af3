"""pyoracle.py -- TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracle libraries.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
only.  Never imported by the uecraytracing_b200 package.
  load("port") -> oracle/liboracle_port.so      (restatement; always buildable)
  load("ref")  -> oracle/_ref/liboracle_ref.so  (the unmodified reference headers; prebuilt where
                                                 /root/reference exists, travels to the GPU box)
  load("best") -> ref when present, else port
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT = os.path.join(HERE, "liboracle_port.so")
REF = os.path.join(HERE, "_ref", "liboracle_ref.so")
REF_BINARY = os.path.join(HERE, "_ref", "raytrace_ref")


def build(ref=True):
    subprocess.check_call(["make", "-C", HERE, "port"] + (["ref"] if ref else []), stdout=subprocess.DEVNULL)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, path):
        self.path = path
        L = self.L = C.CDLL(path)
        vp, u32, u64, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_double
        L.orc_kind.restype = C.c_int
        L.orc_mt19937_nth.restype = u32; L.orc_mt19937_nth.argtypes = [u32, u32]
        L.orc_uniform_nth.restype = dbl; L.orc_uniform_nth.argtypes = [u32, dbl, dbl, u32]
        L.orc_sqrt.restype = dbl; L.orc_sqrt.argtypes = [dbl]
        L.orc_hits.argtypes = [vp, u64, vp, vp, dbl, dbl, vp, vp, vp, vp]
        L.orc_camera_ray.argtypes = [vp, dbl, dbl, dbl, dbl, vp, vp]
        L.orc_scatter.argtypes = [u32, vp, vp, vp, vp, C.c_int, vp, u32, vp, vp, vp, vp]
        L.orc_ray_color.argtypes = [vp, vp, vp, u32, u32, vp, vp]
        L.orc_render_rows.argtypes = [vp, vp, u32, u32, u32, u32, u32, u32, u32, vp, vp]
        L.orc_to_color3b.argtypes = [vp, u64, u32, vp]; L.orc_to_color3b.restype = None
        L.orc_ref4_tuple_hits.argtypes = [u64, vp, vp, dbl, dbl, vp, vp, vp, vp]
        L.orc_ref4_tuple_ray_color.argtypes = [vp, vp, u32, u32, vp]
        self.kind = "reference" if L.orc_kind() == 1 else "port"
        if self.kind == "reference":
            L.orc_ref_default_camera_ray.argtypes = [dbl, dbl, vp, vp]

    # --- primitives
    def mt19937(self, seed, n=0):
        return int(self.L.orc_mt19937_nth(seed, n))

    def uniform(self, seed, a=0.0, b=1.0, n=0):
        return float(self.L.orc_uniform_nth(seed, a, b, n))

    def sqrt(self, s):
        return float(self.L.orc_sqrt(s))

    # --- hits: scene is uecraytracing_b200.Scene (float arrays), rays are float64 arrays
    def hits(self, scene, o, d, t_min=0.001, t_max=float("inf")):
        o = np.ascontiguousarray(o, np.float64).reshape(-1, 3); d = np.ascontiguousarray(d, np.float64).reshape(-1, 3)
        n = len(o)
        t = np.zeros(n); nrm = np.zeros((n, 3)); idx = np.zeros(n, np.int32); ff = np.zeros(n, np.uint8)
        sc = scene.struct()
        assert self.L.orc_hits(C.byref(sc), n, _ptr(o), _ptr(d), t_min, t_max, _ptr(t), _ptr(nrm), _ptr(idx), _ptr(ff)) == 0
        return t, nrm, idx, ff

    def ref4_tuple_hits(self, o, d, t_min=0.001, t_max=float("inf")):
        o = np.ascontiguousarray(o, np.float64).reshape(-1, 3); d = np.ascontiguousarray(d, np.float64).reshape(-1, 3)
        n = len(o)
        t = np.zeros(n); nrm = np.zeros((n, 3)); idx = np.zeros(n, np.int32); ff = np.zeros(n, np.uint8)
        rc = self.L.orc_ref4_tuple_hits(n, _ptr(o), _ptr(d), t_min, t_max, _ptr(t), _ptr(nrm), _ptr(idx), _ptr(ff))
        if rc != 0:
            raise RuntimeError("tuple-typed reference world needs the reference harness (oracle/_ref)")
        return t, nrm, idx, ff

    def ref4_tuple_ray_color(self, o, d, depth, seed):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64); rgb = np.zeros(3)
        if self.L.orc_ref4_tuple_ray_color(_ptr(o), _ptr(d), depth, seed, _ptr(rgb)) != 0:
            raise RuntimeError("needs the reference harness (oracle/_ref)")
        return rgb

    def ref_default_camera_ray(self, u, v):
        o = np.zeros(3); d = np.zeros(3)
        self.L.orc_ref_default_camera_ray(u, v, _ptr(o), _ptr(d))
        return o, d

    def camera_ray(self, camera, s, t, rdx=0.0, rdy=0.0):
        o = np.zeros(3); d = np.zeros(3); cs = camera.struct()
        self.L.orc_camera_ray(C.byref(cs), s, t, rdx, rdy, _ptr(o), _ptr(d))
        return o, d

    def scatter(self, mat_type, mat_param4, d_in, p, normal, front_face, uniforms):
        mp = (C.c_float * 4)(*[float(x) for x in mat_param4])
        d_in = np.ascontiguousarray(d_in, np.float64); p = np.ascontiguousarray(p, np.float64)
        nr = np.ascontiguousarray(normal, np.float64); un = np.ascontiguousarray(uniforms, np.float64)
        ok = C.c_int(0); att = np.zeros(3); oo = np.zeros(3); do = np.zeros(3)
        self.L.orc_scatter(int(mat_type), mp, _ptr(d_in), _ptr(p), _ptr(nr), int(front_face), _ptr(un), len(un),
                           C.byref(ok), _ptr(att), _ptr(oo), _ptr(do))
        return ok.value, att, oo, do

    def ray_color(self, scene, o, d, depth, seed):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64)
        rgb = np.zeros(3); rays = C.c_uint64(0); sc = scene.struct()
        self.L.orc_ray_color(C.byref(sc), _ptr(o), _ptr(d), depth, seed, _ptr(rgb), C.byref(rays))
        return rgb, rays.value

    def render_rows(self, scene, camera, width, height, spp, max_depth, seed, y0, y1):
        out = np.zeros((y1 - y0, width, 3)); rays = C.c_uint64(0)
        sc, cs = scene.struct(), camera.struct()
        self.L.orc_render_rows(C.byref(sc), C.byref(cs), width, height, spp, max_depth, seed, y0, y1, _ptr(out), C.byref(rays))
        return out, rays.value

    def render(self, scene, camera, width, height, spp, max_depth, seed, threads=None):
        """Whole frame, rows sharded over host threads (ctypes releases the GIL)."""
        from concurrent.futures import ThreadPoolExecutor
        threads = threads or os.cpu_count() or 1
        bounds = np.linspace(0, height, min(height, threads * 4) + 1).astype(int)
        jobs = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda ab: self.render_rows(scene, camera, width, height, spp, max_depth, seed, *ab), jobs))
        return np.concatenate([p[0] for p in parts], axis=0), sum(p[1] for p in parts)

    def to_color3b(self, sum_rgb, spp):
        s = np.ascontiguousarray(sum_rgb, np.float64); out = np.zeros(s.shape, np.uint8)
        self.L.orc_to_color3b(_ptr(s), s.size // 3, spp, _ptr(out))
        return out


# --- scenes / cameras without the product library -------------------------------------------------
# bench.py's reference arm must not map libuecrt.so (the driver records which .so files each arm loads), so it
# takes its scene and camera arrays from tests/golden/bench_scenes.npz (written by tools/make_bench_scenes.py
# from the product's host scene builders) through these two stand-ins for uecraytracing_b200.Scene / Camera:
# same .struct() layout as include/uecrt.h uecrt_scene / uecrt_camera.
class _SceneStruct(C.Structure):
    _fields_ = [("n", C.c_uint32), ("geom", C.c_void_p), ("mat_type", C.c_void_p), ("mat_param", C.c_void_p)]


class _CameraStruct(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("lens_radius", C.c_float)]


class PlainScene:
    def __init__(self, geom, mat_type, mat_param):
        self.geom = np.ascontiguousarray(geom, np.float32).reshape(-1, 4)
        self.mat_type = np.ascontiguousarray(mat_type, np.uint32).reshape(-1)
        self.mat_param = np.ascontiguousarray(mat_param, np.float32).reshape(-1, 4)
        self.n = len(self.geom)

    def struct(self):
        return _SceneStruct(self.n, _ptr(self.geom).value, _ptr(self.mat_type).value, _ptr(self.mat_param).value)


class PlainCamera:
    FIELDS = ("origin", "lower_left", "horizontal", "vertical", "u", "v")

    def __init__(self, vec18, lens_radius):
        self.s = _CameraStruct()
        v = np.asarray(vec18, np.float32).reshape(6, 3)
        for k, name in enumerate(self.FIELDS):
            setattr(self.s, name, (C.c_float * 3)(*[float(x) for x in v[k]]))
        self.s.lens_radius = float(lens_radius)

    def struct(self):
        return self.s


BENCH_SCENES = os.path.join(os.path.dirname(HERE), "tests", "golden", "bench_scenes.npz")


def bench_scene(name, camera):
    """(PlainScene, PlainCamera) of a bench.py workload from the committed fixture (no product code involved)."""
    z = np.load(BENCH_SCENES)
    sc = PlainScene(z[f"{name}_geom"], z[f"{name}_mat_type"], z[f"{name}_mat_param"])
    cam = PlainCamera(z[f"cam_{camera}_vec"], float(z[f"cam_{camera}_lens"]))
    return sc, cam


def load(which="best"):
    if which in ("ref", "best") and os.path.exists(REF):
        return Oracle(REF)
    if which == "ref":
        raise FileNotFoundError(REF)
    if not os.path.exists(PORT):
        build(ref=False)
    return Oracle(PORT)
