"""Python-side plumbing over the C ABI: numpy in, numpy out (host entry points), raw device
pointers for the device-resident entry points (bench.py passes torch tensors' data_ptr())."""
import ctypes as C

import numpy as np

from ._lib import (CameraStruct, ParamsStruct, SceneStruct, StatsStruct, UecrtError, check, lib,
                   FLAG_EXACT_HITS, FLAG_SCAN_FP32, FLAG_SCAN_TENSOR, FLAG_SCAN_GRID, SCAN_AUTO, SCAN_FP32, SCAN_TENSOR, SCAN_GRID)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def image_height(width):
    """uint32(width / (16/9)) -- source.cpp:59-62."""
    return int(lib.uecrt_host_image_height(int(width)))


def device_count():
    return int(lib.uecrt_device_count())


class Scene:
    """Flattened hittable_list: geom (n,4) f32 {cx,cy,cz,r}, mat_type (n,) u32, mat_param (n,4) f32."""

    def __init__(self, geom, mat_type, mat_param):
        self.geom = np.ascontiguousarray(geom, dtype=np.float32).reshape(-1, 4)
        self.mat_type = np.ascontiguousarray(mat_type, dtype=np.uint32).reshape(-1)
        self.mat_param = np.ascontiguousarray(mat_param, dtype=np.float32).reshape(-1, 4)
        if not (len(self.geom) == len(self.mat_type) == len(self.mat_param)):
            raise ValueError("scene arrays disagree in length")

    @property
    def n(self):
        return len(self.geom)

    def struct(self):
        return SceneStruct(self.n, _ptr(self.geom).value, _ptr(self.mat_type).value, _ptr(self.mat_param).value)

    @classmethod
    def named(cls, name, seed=1, n=10000):
        """'ref4' | 'ref4d' | 'final' | 'dense' -- built by the C++ host code (host/yk/scenes.hpp)."""
        cnt = lib.uecrt_host_scene_count(name.encode(), n, seed)
        if cnt == 0 and not (name == "dense" and n == 0):
            raise ValueError(f"unknown scene {name!r}")
        geom = np.zeros((cnt, 4), np.float32)
        mt = np.zeros(cnt, np.uint32)
        mp = np.zeros((cnt, 4), np.float32)
        if cnt:
            check(lib.uecrt_host_scene_fill(name.encode(), n, seed, cnt, _ptr(geom), _ptr(mt), _ptr(mp)), "scene_fill")
        return cls(geom, mt, mp)


def load_scene_file(path):
    """(Scene, Camera or None) from a text scene description (host/yk/scene_io.hpp)."""
    n = lib.uecrt_host_scene_file_count(str(path).encode())
    if n < 0:
        raise ValueError(lib.uecrt_host_last_error().decode(errors="replace"))
    geom = np.zeros((n, 4), np.float32); mt = np.zeros(n, np.uint32); mp = np.zeros((n, 4), np.float32)
    cam = CameraStruct(); has = C.c_int(0)
    check(lib.uecrt_host_scene_file_fill(str(path).encode(), n, _ptr(geom), _ptr(mt), _ptr(mp), C.byref(cam), C.byref(has)),
          "scene_file_fill")
    return Scene(geom, mt, mp), (Camera(cam) if has.value else None)


def save_scene_file(path, scene):
    check(lib.uecrt_host_scene_file_write(str(path).encode(), scene.n, _ptr(scene.geom), _ptr(scene.mat_type),
                                          _ptr(scene.mat_param)), "scene_file_write")


def write_pfm(path, accum, spp):
    a = np.ascontiguousarray(accum, np.float32)
    h, w, _ = a.shape
    check(lib.uecrt_host_write_pfm(str(path).encode(), w, h, _ptr(a), int(spp)), "uecrt_host_write_pfm")


class Camera:
    def __init__(self, struct):
        self.s = struct

    @classmethod
    def named(cls, name):
        s = CameraStruct()
        check(lib.uecrt_host_camera(name.encode(), C.byref(s)), "camera")
        return cls(s)

    @classmethod
    def lookat(cls, lookfrom, lookat, vup=(0, 1, 0), vfov=20.0, aspect=16.0 / 9.0, aperture=0.0, focus_dist=1.0):
        s = CameraStruct()
        f = (C.c_double * 3)(*lookfrom); a = (C.c_double * 3)(*lookat); v = (C.c_double * 3)(*vup)
        check(lib.uecrt_host_camera_lookat(f, a, v, vfov, aspect, aperture, focus_dist, C.byref(s)), "camera_lookat")
        return cls(s)

    def struct(self):
        return self.s


def stats_to_dict(s):
    return {k: getattr(s, k) for k, _ in StatsStruct._fields_}


def _params(width, height, spp, max_depth, seed, sample_begin=0, sample_count=0, num_gpus=1, flags=0, shard=(0, 0)):
    return ParamsStruct(width, height, spp, max_depth, seed, sample_begin, sample_count, num_gpus, flags, shard[0], shard[1])


_SCAN_FLAGS = {None: 0, "auto": 0, "fp32": FLAG_SCAN_FP32, "tensor": FLAG_SCAN_TENSOR, "grid": FLAG_SCAN_GRID}


def render(scene, camera, width=400, spp=100, max_depth=50, seed=0, gpus=1, exact=False, return_accum=False,
           sample_begin=0, sample_count=0, scan=None, shard=(0, 0)):
    """uecrt_render with HOST buffers: the call the re-hosted yk::render() makes."""
    height = image_height(width)
    rgb = np.zeros((height, width, 3), np.uint8)
    accum = np.zeros((height, width, 4), np.float32) if return_accum else None
    st = StatsStruct()
    sc, cam = scene.struct(), camera.struct()
    p = _params(width, height, spp, max_depth, seed, sample_begin, sample_count, gpus,
                (FLAG_EXACT_HITS if exact else 0) | _SCAN_FLAGS[scan], shard)
    check(lib.uecrt_render(C.byref(sc), C.byref(cam), C.byref(p), _ptr(rgb), _ptr(accum) if return_accum else None,
                           C.byref(st)), "uecrt_render")
    return rgb, accum, stats_to_dict(st)


class FrameQueue:
    """uecrt_queue_*: K1 of frame k + 1 overlaps the D2H (pinned, copy engine) and the host-side handling of frame k."""

    def __init__(self, scene, camera, width):
        self.width, self.height = width, image_height(width)
        self.q = C.c_void_p()
        sc, cam = scene.struct(), camera.struct()
        check(lib.uecrt_queue_create(C.byref(sc), C.byref(cam), self.width, self.height, C.byref(self.q)), "uecrt_queue_create")

    def submit(self, frame_id, spp=100, max_depth=50, seed=0, camera=None, exact=False, scan=None):
        p = _params(self.width, self.height, spp, max_depth, seed, 0, 0, 1, (FLAG_EXACT_HITS if exact else 0) | _SCAN_FLAGS[scan])
        cam = camera.struct() if camera is not None else None
        check(lib.uecrt_queue_submit(self.q, C.byref(cam) if cam is not None else None, C.byref(p), int(frame_id)), "uecrt_queue_submit")

    def next(self):
        """(frame_id, rgb copy (H, W, 3) uint8, K1 ms) of the oldest frame in flight; blocks until its RGB has landed."""
        ptr, fid, ms = C.c_void_p(), C.c_uint64(), C.c_float()
        check(lib.uecrt_queue_next(self.q, C.byref(ptr), C.byref(fid), C.byref(ms)), "uecrt_queue_next")
        buf = (C.c_uint8 * (self.width * self.height * 3)).from_address(ptr.value)
        return int(fid.value), np.frombuffer(buf, np.uint8).reshape(self.height, self.width, 3).copy(), float(ms.value)

    def close(self):
        if self.q:
            lib.uecrt_queue_destroy(self.q)
            self.q = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def primary_hits(scene, o, d, t_min=0.001, t_max=float("inf"), exact=True):
    o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    n = len(o)
    t = np.zeros(n, np.float64); nrm = np.zeros((n, 3), np.float64)
    idx = np.zeros(n, np.int32); ff = np.zeros(n, np.uint8)
    sc = scene.struct()
    check(lib.uecrt_primary_hits(C.byref(sc), n, _ptr(o), _ptr(d), t_min, t_max, 1 if exact else 0, _ptr(t), _ptr(nrm),
                                 _ptr(idx), _ptr(ff)), "uecrt_primary_hits")
    return t, nrm, idx, ff


def camera_rays(camera, width, seed, pixel, sample):
    """(origins, unnormalised directions) of the camera rays K1 starts paths (pixel, sample) with (uecrt_debug_camera_rays)."""
    pixel = np.ascontiguousarray(pixel, np.uint32); sample = np.ascontiguousarray(sample, np.uint32)
    n = len(pixel)
    o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32)
    cam = camera.struct()
    check(lib.uecrt_debug_camera_rays(C.byref(cam), width, image_height(width), int(seed), n, _ptr(pixel), _ptr(sample), _ptr(o), _ptr(d)),
          "uecrt_debug_camera_rays")
    return o, d


def tensor_candidates(scene, o, d):
    """(flags (n_rays, n_spheres) uint8, spheres_per_flag): the tensor-core scan's verdict alone (uecrt_debug_tensor_candidates)."""
    o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    flags = np.zeros((len(o), scene.n), np.uint8)
    q = C.c_int32(0)
    sc = scene.struct()
    check(lib.uecrt_debug_tensor_candidates(C.byref(sc), len(o), _ptr(o), _ptr(d), _ptr(flags), C.byref(q)), "uecrt_debug_tensor_candidates")
    return flags, q.value


def scatter(mat_type, mat_param, d_in, p, normal, front_face, uniforms4):
    mt = np.ascontiguousarray(mat_type, np.uint32); n = len(mt)
    mp = np.ascontiguousarray(mat_param, np.float32).reshape(n, 4)
    d_in = np.ascontiguousarray(d_in, np.float32).reshape(n, 3); p = np.ascontiguousarray(p, np.float32).reshape(n, 3)
    nr = np.ascontiguousarray(normal, np.float32).reshape(n, 3); ff = np.ascontiguousarray(front_face, np.uint8)
    un = np.ascontiguousarray(uniforms4, np.float32).reshape(n, 4)
    ok = np.zeros(n, np.uint8); att = np.zeros((n, 3), np.float32)
    oo = np.zeros((n, 3), np.float32); do = np.zeros((n, 3), np.float32)
    check(lib.uecrt_scatter(n, _ptr(mt), _ptr(mp), _ptr(d_in), _ptr(p), _ptr(nr), _ptr(ff), _ptr(un), _ptr(ok), _ptr(att),
                            _ptr(oo), _ptr(do)), "uecrt_scatter")
    return ok, att, oo, do


def philox(seed, counters4):
    c = np.ascontiguousarray(counters4, np.uint32).reshape(-1, 4)
    out = np.zeros_like(c)
    check(lib.uecrt_philox(len(c), seed, _ptr(c), _ptr(out)), "uecrt_philox")
    return out


class Context:
    """Device-resident scene on the current CUDA device (uecrt_ctx)."""

    def __init__(self, scene, camera):
        self._h = C.c_void_p()
        sc, cam = scene.struct(), camera.struct()
        check(lib.uecrt_ctx_create(C.byref(sc), C.byref(cam), C.byref(self._h)), "uecrt_ctx_create")

    def set_scene(self, scene, camera):
        """Host->device re-upload of scene and camera into this context (no allocation when it fits)."""
        sc, cam = scene.struct(), camera.struct()
        check(lib.uecrt_ctx_set_scene(self._h, C.byref(sc), C.byref(cam)), "uecrt_ctx_set_scene")

    def accumulate(self, d_accum_ptr, width, height, spp, max_depth, seed, sample_begin=0, sample_count=0, stream=0,
                   flags=0, shard=(0, 0)):
        """shard = (index, count): this call renders the index-th of count equal ranges of the frame's work items."""
        p = _params(width, height, spp, max_depth, seed, sample_begin, sample_count, 1, flags, shard)
        check(lib.uecrt_accumulate_device(self._h, C.byref(p), C.c_void_p(d_accum_ptr), C.c_void_p(stream), None),
              "uecrt_accumulate_device")

    def primary_hits(self, o, d, t_min=0.001, t_max=float("inf"), exact=True):
        """Closest hits of caller rays against this context's resident scene (uecrt_ctx_primary_hits)."""
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = len(o)
        t = np.zeros(n, np.float64); nrm = np.zeros((n, 3), np.float64)
        idx = np.zeros(n, np.int32); ff = np.zeros(n, np.uint8)
        check(lib.uecrt_ctx_primary_hits(self._h, n, _ptr(o), _ptr(d), t_min, t_max, 1 if exact else 0, _ptr(t), _ptr(nrm),
                                         _ptr(idx), _ptr(ff)), "uecrt_ctx_primary_hits")
        return t, nrm, idx, ff

    def stats(self):
        st = StatsStruct()
        check(lib.uecrt_ctx_stats(self._h, C.byref(st)), "uecrt_ctx_stats")
        return stats_to_dict(st)

    def close(self):
        if self._h:
            lib.uecrt_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def resolve_device(d_accum_ptr, width, height, spp, d_rgb_ptr, stream=0):
    check(lib.uecrt_resolve_device(C.c_void_p(d_accum_ptr), width, height, spp, C.c_void_p(d_rgb_ptr), C.c_void_p(stream)),
          "uecrt_resolve_device")


def write_png(path, rgb):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    check(lib.uecrt_host_write_png(str(path).encode(), w, h, _ptr(rgb)), "uecrt_host_write_png")


def philox_host(seed, counter4):
    """Philox4x32-10 block computed by the host build of csrc/philox.cuh (no GPU needed)."""
    c = (C.c_uint32 * 4)(*[int(x) for x in counter4]); out = (C.c_uint32 * 4)()
    lib.uecrt_host_philox(seed, c, out)
    return [int(x) for x in out]


def set_tile_spheres(n):
    """0 = automatic residency/tiling; n > 0 forces shared-memory tiles of n spheres (see include/uecrt.h)."""
    lib.uecrt_set_tile_spheres(int(n))


def set_rays_per_thread(n):
    """Paths per K1 thread sharing one scene scan (1 or 2); applies to scenes loaded afterwards."""
    lib.uecrt_set_rays_per_thread(int(n))


def set_scan_mode(mode):
    """'auto' | 'fp32' | 'tensor' | 'grid': process default of the candidate scan (include/uecrt.h UECRT_SCAN_*)."""
    lib.uecrt_set_scan_mode({"auto": SCAN_AUTO, "fp32": SCAN_FP32, "tensor": SCAN_TENSOR, "grid": SCAN_GRID}[mode])


def set_tensor_pipeline(name):
    """'self' | 'scanwarps': K1 pipeline of the resident tensor-core scan (include/uecrt.h UECRT_TENSOR_PIPELINE_*)."""
    lib.uecrt_set_tensor_pipeline({"scanwarps": 0, "self": 1}[name])


def tensor_scan_max_spheres():
    return int(lib.uecrt_tensor_scan_max_spheres())


def tensor_scan_resident_spheres():
    """Largest scene whose tensor-scan operand stays resident in shared memory; larger scenes are swept in slices."""
    return int(lib.uecrt_tensor_scan_resident_spheres())
