/* uecrt_host.h -- C entry points to the HOST side of the path (scene / camera builders, image
 * geometry, PNG output): the callers and data formats either side of the hot path.  They are
 * implemented by the yk:: host headers (uecraytracing_b200/host/yk) and exported from
 * libuecrt.so so that Python (tests, bench.py) and the raytrace CLI build byte-identical
 * scenes.  None of them touches the GPU. */
#ifndef UECRT_HOST_H
#define UECRT_HOST_H
#include "uecrt.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Scenes: "ref4" (source.cpp:103-112), "ref4d" (id2 -> dielectric), "final" (RTIOW final
 * random scene, `seed` drives std::mt19937), "dense" (n_hint spheres, BASELINE config 4).
 * uecrt_host_scene_count returns the number of spheres (0 for an unknown name);
 * uecrt_host_scene_fill writes up to `cap` of them and returns 0 on success. */
uint32_t uecrt_host_scene_count(const char* name, uint32_t n_hint, uint32_t seed);
int uecrt_host_scene_fill(const char* name, uint32_t n_hint, uint32_t seed, uint32_t cap,
                          uecrt_float4* geom, uint32_t* mat_type, uecrt_float4* mat_param);

/* Scene description files (host/yk/scene_io.hpp: `sphere cx cy cz r lambertian|metal|dielectric ...`,
 * optional `camera ...`).  uecrt_host_scene_file_count returns the sphere count or -1 on a parse error
 * (message through uecrt_host_last_error); _fill copies the arrays and, when the file has a camera
 * line, the camera (has_camera_out = 1).  uecrt_host_scene_file_write writes arrays back as text. */
int32_t uecrt_host_scene_file_count(const char* path);
int uecrt_host_scene_file_fill(const char* path, uint32_t cap, uecrt_float4* geom, uint32_t* mat_type,
                               uecrt_float4* mat_param, uecrt_camera* cam_out, int* has_camera_out);
int uecrt_host_scene_file_write(const char* path, uint32_t n, const uecrt_float4* geom,
                                const uint32_t* mat_type, const uecrt_float4* mat_param);
const char* uecrt_host_last_error(void);

/* Cameras: "ref" = yk::camera<T>{} (camera.hpp:16-27); "final" / "dense" = RTIOW look-at. */
int uecrt_host_camera(const char* name, uecrt_camera* out);
int uecrt_host_camera_lookat(const double from[3], const double at[3], const double vup[3],
                             double vfov_deg, double aspect, double aperture, double focus_dist,
                             uecrt_camera* out);

/* image_height = uint32(width / (16/9)) (source.cpp:59-62) */
uint32_t uecrt_host_image_height(uint32_t width);

/* Portable float map (PFM, "PF", little-endian, bottom row first) of the linear image sum/spp taken
 * from a W*H*4 accumulation buffer.  Returns 0 on success. */
int uecrt_host_write_pfm(const char* path, uint32_t width, uint32_t height, const float* accum4, uint32_t spp);

/* 8-bit RGB PNG, stride 3*width, top row first (what source.cpp:224-226 asks of
 * stbi_write_png).  Returns 0 on success. */
int uecrt_host_write_png(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb);

/* Philox4x32-10 on the host (same code as the device, csrc/philox.cuh), for known-answer tests */
void uecrt_host_philox(uint64_t seed, const uint32_t counter[4], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
