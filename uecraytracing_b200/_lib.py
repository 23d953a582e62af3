"""ctypes declarations for include/uecrt.h and include/uecrt_host.h."""
import ctypes as C
import os

MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC = 0, 1, 2
FLAG_EXACT_HITS, FLAG_NO_COUNTERS, FLAG_SCAN_FP32, FLAG_SCAN_TENSOR, FLAG_SCAN_GRID = 1, 2, 4, 8, 16
SCAN_AUTO, SCAN_FP32, SCAN_TENSOR, SCAN_GRID = 0, 1, 2, 3


class UecrtError(RuntimeError):
    pass


class Float4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class SceneStruct(C.Structure):
    _fields_ = [("n", C.c_uint32), ("geom", C.c_void_p), ("mat_type", C.c_void_p), ("mat_param", C.c_void_p)]


class CameraStruct(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("lens_radius", C.c_float)]


class ParamsStruct(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("max_depth", C.c_uint32),
                ("seed", C.c_uint64), ("sample_begin", C.c_uint32), ("sample_count", C.c_uint32),
                ("num_gpus", C.c_int32), ("flags", C.c_uint32), ("shard_index", C.c_uint32), ("shard_count", C.c_uint32)]


class StatsStruct(C.Structure):
    _fields_ = [("rays_traced", C.c_uint64), ("sphere_tests", C.c_uint64), ("paths", C.c_uint64),
                ("kernel_ms", C.c_double), ("reduce_ms", C.c_double), ("resolve_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("sm_count", C.c_int32), ("sm_clock_khz", C.c_int32),
                ("k1_blocks", C.c_int32), ("k1_threads", C.c_int32), ("k1_smem_bytes", C.c_int32), ("k1_regs", C.c_int32),
                ("tiles", C.c_int32), ("tile_spheres", C.c_int32), ("scan_mode", C.c_int32), ("reserved_", C.c_int32)]


# UECRT_LIB selects an alternative build of the same ABI (kernel tuning experiments only)
lib_path = os.environ.get("UECRT_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libuecrt.so")

# every symbol include/uecrt.h and include/uecrt_host.h declare
EXPORTS = [
    "uecrt_last_error", "uecrt_version", "uecrt_device_count", "uecrt_render", "uecrt_release", "uecrt_set_tile_spheres", "uecrt_set_rays_per_thread", "uecrt_set_scan_mode", "uecrt_set_tensor_pipeline", "uecrt_tensor_scan_max_spheres", "uecrt_tensor_scan_resident_spheres", "uecrt_primary_hits", "uecrt_scatter", "uecrt_debug_tensor_candidates", "uecrt_debug_camera_rays",
    "uecrt_philox", "uecrt_ctx_create", "uecrt_ctx_destroy", "uecrt_ctx_set_scene", "uecrt_accumulate_device", "uecrt_resolve_device",
    "uecrt_ctx_stats", "uecrt_ctx_primary_hits", "uecrt_queue_create", "uecrt_queue_destroy", "uecrt_queue_submit", "uecrt_queue_next", "uecrt_host_scene_count", "uecrt_host_scene_fill", "uecrt_host_camera",
    "uecrt_host_camera_lookat", "uecrt_host_image_height", "uecrt_host_write_png", "uecrt_host_philox", "uecrt_host_scene_file_count", "uecrt_host_scene_file_fill",
    "uecrt_host_scene_file_write", "uecrt_host_last_error", "uecrt_host_write_pfm",
]


def _load():
    if not os.path.exists(lib_path):
        raise UecrtError(f"{lib_path} is missing: build it with `make` (or __graft_entry__.build()); "
                         "there is no Python/CPU fallback for the CUDA path")
    L = C.CDLL(lib_path)
    vp, u32, u64, i32, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_double
    L.uecrt_last_error.restype = C.c_char_p
    L.uecrt_version.restype = C.c_int
    L.uecrt_device_count.restype = C.c_int
    L.uecrt_render.argtypes = [vp, vp, vp, vp, vp, vp]
    L.uecrt_release.restype = None
    L.uecrt_set_tile_spheres.argtypes = [u32]
    L.uecrt_set_tile_spheres.restype = None
    L.uecrt_set_rays_per_thread.argtypes = [C.c_int]
    L.uecrt_set_rays_per_thread.restype = None
    L.uecrt_set_scan_mode.argtypes = [C.c_int]
    L.uecrt_set_scan_mode.restype = None
    L.uecrt_set_tensor_pipeline.argtypes = [C.c_int]
    L.uecrt_set_tensor_pipeline.restype = None
    L.uecrt_tensor_scan_max_spheres.restype = C.c_int
    L.uecrt_tensor_scan_resident_spheres.restype = C.c_int
    L.uecrt_primary_hits.argtypes = [vp, u64, vp, vp, dbl, dbl, C.c_int, vp, vp, vp, vp]
    L.uecrt_debug_tensor_candidates.argtypes = [vp, u64, vp, vp, vp, vp]
    L.uecrt_debug_tensor_candidates.restype = C.c_int
    L.uecrt_debug_camera_rays.argtypes = [vp, u32, u32, u64, u64, vp, vp, vp, vp]
    L.uecrt_debug_camera_rays.restype = C.c_int
    L.uecrt_scatter.argtypes = [u64] + [vp] * 11
    L.uecrt_philox.argtypes = [u64, u64, vp, vp]
    L.uecrt_ctx_create.argtypes = [vp, vp, C.POINTER(vp)]
    L.uecrt_ctx_destroy.argtypes = [vp]
    L.uecrt_ctx_set_scene.argtypes = [vp, vp, vp]
    L.uecrt_ctx_set_scene.restype = C.c_int
    L.uecrt_ctx_destroy.restype = None
    L.uecrt_accumulate_device.argtypes = [vp, vp, vp, vp, vp]
    L.uecrt_resolve_device.argtypes = [vp, u32, u32, u32, vp, vp]
    L.uecrt_ctx_stats.argtypes = [vp, vp]
    L.uecrt_ctx_primary_hits.argtypes = [vp, u64, vp, vp, dbl, dbl, C.c_int, vp, vp, vp, vp]
    L.uecrt_ctx_primary_hits.restype = C.c_int
    L.uecrt_queue_create.argtypes = [vp, vp, u32, u32, C.POINTER(vp)]
    L.uecrt_queue_create.restype = C.c_int
    L.uecrt_queue_destroy.argtypes = [vp]
    L.uecrt_queue_destroy.restype = None
    L.uecrt_queue_submit.argtypes = [vp, vp, vp, u64]
    L.uecrt_queue_submit.restype = C.c_int
    L.uecrt_queue_next.argtypes = [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(C.c_float)]
    L.uecrt_queue_next.restype = C.c_int
    L.uecrt_host_scene_count.argtypes = [C.c_char_p, u32, u32]
    L.uecrt_host_scene_count.restype = u32
    L.uecrt_host_scene_fill.argtypes = [C.c_char_p, u32, u32, u32, vp, vp, vp]
    L.uecrt_host_camera.argtypes = [C.c_char_p, vp]
    L.uecrt_host_camera_lookat.argtypes = [vp, vp, vp, dbl, dbl, dbl, dbl, vp]
    L.uecrt_host_image_height.argtypes = [u32]
    L.uecrt_host_image_height.restype = u32
    L.uecrt_host_write_png.argtypes = [C.c_char_p, u32, u32, vp]
    L.uecrt_host_philox.argtypes = [u64, vp, vp]
    L.uecrt_host_scene_file_count.argtypes = [C.c_char_p]
    L.uecrt_host_scene_file_count.restype = i32
    L.uecrt_host_scene_file_fill.argtypes = [C.c_char_p, u32, vp, vp, vp, vp, vp]
    L.uecrt_host_scene_file_fill.restype = C.c_int
    L.uecrt_host_scene_file_write.argtypes = [C.c_char_p, u32, vp, vp, vp]
    L.uecrt_host_scene_file_write.restype = C.c_int
    L.uecrt_host_last_error.restype = C.c_char_p
    L.uecrt_host_write_pfm.argtypes = [C.c_char_p, u32, u32, vp, u32]
    L.uecrt_host_write_pfm.restype = C.c_int
    L.uecrt_host_philox.restype = None
    for name in ("uecrt_render", "uecrt_primary_hits", "uecrt_scatter", "uecrt_philox", "uecrt_ctx_create",
                 "uecrt_accumulate_device", "uecrt_resolve_device", "uecrt_ctx_stats", "uecrt_host_scene_fill",
                 "uecrt_host_camera", "uecrt_host_camera_lookat", "uecrt_host_write_png"):
        getattr(L, name).restype = C.c_int
    return L


lib = _load()


def check(rc, what):
    if rc != 0:
        raise UecrtError(f"{what} failed (code {rc}): {lib.uecrt_last_error().decode(errors='replace')}")
