"""uecraytracing_b200 -- B200-native path-tracing hot path of yaito3014/UECRayTracing.

Thin ctypes binding over the C ABI in ``include/uecrt.h`` (``libuecrt.so``: hand-written sm_100a
CUDA kernels).  Python is plumbing for tests and ``bench.py``; the product is the shared library
and the ``bin/raytrace`` drop-in CLI.  There is no CPU fallback: importing works without a GPU
(scene/camera builders are host code), every compute call raises ``UecrtError`` without one.
"""
from ._lib import (  # noqa: F401
    UecrtError, lib, lib_path, Float4, SceneStruct, CameraStruct, ParamsStruct, StatsStruct,
    MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, FLAG_EXACT_HITS, FLAG_NO_COUNTERS,
    FLAG_SCAN_FP32, FLAG_SCAN_TENSOR, FLAG_SCAN_GRID,
)
from .api import (  # noqa: F401
    Scene, Camera, Context, FrameQueue, image_height, render, primary_hits, scatter, philox, resolve_device,
    write_png, device_count, stats_to_dict, philox_host, set_tile_spheres, set_rays_per_thread, load_scene_file, save_scene_file, write_pfm,
    set_scan_mode, set_tensor_pipeline, tensor_scan_max_spheres, tensor_scan_resident_spheres, tensor_candidates, camera_rays,
)

__all__ = [
    "UecrtError", "lib", "lib_path", "Scene", "Camera", "Context", "FrameQueue", "image_height", "render",
    "primary_hits", "scatter", "philox", "resolve_device", "write_png", "device_count",
    "stats_to_dict", "MAT_LAMBERTIAN", "MAT_METAL", "MAT_DIELECTRIC", "FLAG_EXACT_HITS",
    "FLAG_NO_COUNTERS", "FLAG_SCAN_FP32", "FLAG_SCAN_TENSOR", "FLAG_SCAN_GRID", "set_scan_mode", "set_tensor_pipeline", "tensor_scan_max_spheres", "tensor_scan_resident_spheres",
]
