/* uecrt.h -- C ABI of the B200-native path-tracing hot path (libuecrt.so).
 *
 * The reference (yaito3014/UECRayTracing) has no plugin / FFI interface: it is one
 * translation unit whose main() calls yk::render() (source.cpp:98-178, 221).  This header
 * is the thin extern "C" layer the re-hosted yk::render() binds instead; every entry point
 * names the reference code it replaces.  Plain pointers and sizes only; no C++ or torch
 * types cross this boundary.  No CPU fallback exists behind it: every compute entry point
 * returns UECRT_ERR_NO_DEVICE when no sm_100 device is usable.
 *
 * Conventions: return 0 = OK, non-zero = failure (uecrt_last_error() gives the text).
 * The caller owns every buffer it passes.  "host" entry points copy to/from the device
 * inside the call; "_device" entry points take device pointers and a cudaStream_t (as
 * void*), so a host runtime that already owns device memory (torch, in bench.py) can keep
 * the data resident.
 */
#ifndef UECRT_H
#define UECRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UECRT_VERSION 2

enum {
  UECRT_OK = 0,
  UECRT_ERR_INVALID = 1,   /* bad argument */
  UECRT_ERR_NO_DEVICE = 2, /* no CUDA device / not sm_100 */
  UECRT_ERR_CUDA = 3,      /* a CUDA call failed */
  UECRT_ERR_NCCL = 4       /* a NCCL call failed (multi-GPU render) */
};

/* material kinds; 0 and 1 are the reference's (yk/material.hpp:44-76), 2 is the RTIOW
 * dielectric named by north_star (absent from the reference). */
enum { UECRT_MAT_LAMBERTIAN = 0, UECRT_MAT_METAL = 1, UECRT_MAT_DIELECTRIC = 2 };

typedef struct { float x, y, z, w; } uecrt_float4;

/* Flattened hittable_list (yk/hittable_list.hpp:18-30) of spheres (yk/sphere.hpp:16-23).
 * Index order == insertion order == hit_record::id (hittable_list.hpp:41); the closest-hit
 * tie rule "later index wins" depends on it.  Host pointers. */
typedef struct {
  uint32_t n;
  const uecrt_float4* geom;      /* {cx, cy, cz, radius}                    */
  const uint32_t* mat_type;      /* UECRT_MAT_*                             */
  const uecrt_float4* mat_param; /* {albedo r,g,b, fuzz (metal) | ior (dielectric)} */
} uecrt_scene;

/* yk::camera (yk/camera.hpp:14-38) extended with the RTIOW look-at / defocus fields.
 * get_ray(s,t) = { origin + off, lower_left + s*horizontal + t*vertical - origin - off },
 * off = lens_radius * (rd.x*u + rd.y*v).  The reference camera is origin 0,
 * lower_left (-16/9,-1,-1), horizontal (32/9,0,0), vertical (0,2,0), lens_radius 0. */
typedef struct {
  float origin[3], lower_left[3], horizontal[3], vertical[3], u[3], v[3];
  float lens_radius;
} uecrt_camera;

typedef struct {
  uint32_t width, height; /* height = uint32(width / (16/9)) in the reference (source.cpp:59-62) */
  uint32_t spp;           /* YK_SPP, source.cpp:47-49       */
  uint32_t max_depth;     /* YK_MAX_DEPTH, source.cpp:51-53 */
  uint64_t seed;          /* Philox key                     */
  uint32_t sample_begin;  /* this call renders sample ids [sample_begin, sample_begin+sample_count) */
  uint32_t sample_count;  /*   of every pixel; 0 means "all spp" (sample_begin ignored)            */
  int32_t num_gpus;       /* uecrt_render only: shard the frame over this many devices (>=1)       */
  uint32_t flags;         /* UECRT_FLAG_*                   */
  /* Work-item sharding (one process per GPU): a frame's work items are its (pixel, sample) paths, numbered
   * sample-major, so a contiguous range of them is "a few whole samples of every pixel plus part of one more".
   * shard_count > 0: this call renders the shard_index-th of shard_count equal ranges of the items selected by
   * [sample_begin, +sample_count); every device of a G-way split gets items/G paths whatever spp is (spp-granular
   * shards differ by a whole sample per pixel: 62 vs 63 spp at 500 spp over 8 GPUs).  The Philox streams are keyed
   * by the global (pixel, sample), so the summed image does not depend on how the items were split.
   * 0 = no sharding.  uecrt_render shards this way over num_gpus devices by itself. */
  uint32_t shard_index, shard_count;
} uecrt_params;

enum {
  UECRT_FLAG_EXACT_HITS = 1u, /* resolve every closest hit in FP64 (reference-exact index) inside K1.
                                 Without it (FAST, the default) the candidates of a ray are compared by
                                 their FP32 roots and only the WINNER's root is solved again in FP64
                                 (csrc/closest_hit.cuh refine_winner).  Stated FAST tolerance against
                                 sphere.hpp:25-48 in double, tested on 810 000 primary and 600 000 random
                                 rays (tests/test_gpu_hits.py, profiles/fast_accuracy_r2.log): t <= 1e-9
                                 relative, normal <= 1e-9 absolute, front_face equal, wherever the same
                                 sphere wins; the winning sphere may differ from the reference's when two
                                 roots lie within ~1e-6 relative of each other (<= 2e-4 of the rays
                                 allowed, 0 observed). */
  UECRT_FLAG_NO_COUNTERS = 2u, /* do not count rays/sphere tests (saves two atomics per CTA)         */
  UECRT_FLAG_SCAN_FP32 = 4u,   /* force the FP32-pipe candidate scan (packed FFMA2 from shared memory) */
  UECRT_FLAG_SCAN_TENSOR = 8u, /* force the tensor-core candidate scan (tcgen05 TF32 x3, TMEM), any
                                  scene size: resident or swept sphere operand, see below             */
  UECRT_FLAG_SCAN_GRID = 16u   /* walk a uniform grid instead of scanning every sphere (opt-in spatial
                                  index: same closest hit, different algorithmic work)                */
};
/* scan used by a render / hit call: the all-pairs discriminant test of hittable_list::hit_impl
 * (hittable_list.hpp:32-58) runs either on the FP32 pipe or on the tensor cores; both produce a
 * superset of the reference's "discriminant >= 0" spheres and share the resolve phase, so the
 * closest hit is the same. */
enum { UECRT_SCAN_AUTO = 0, UECRT_SCAN_FP32 = 1, UECRT_SCAN_TENSOR = 2, UECRT_SCAN_GRID = 3 };
/* UECRT_SCAN_GRID replaces "test every object" (hittable_list.hpp:32-58) by a uniform grid over the spheres:
 * a ray tests the spheres listed in the cells it passes (3-D DDA) plus the spheres too large for the
 * grid, with the same per-sphere test and the same tie rule, so the closest hit is the same sphere.  It
 * is never chosen automatically: the brute-force scans are the path north_star measures, and the
 * grid's work is not 18 FLOP x N x rays.  uecrt_stats.sphere_tests counts the tests actually executed. */

typedef struct {
  uint64_t rays_traced;  /* closest-hit scans executed (one per path segment)          */
  uint64_t sphere_tests; /* rays_traced * n (brute-force scans); executed tests (grid)  */
  uint64_t paths;        /* (pixel,sample) paths finished                              */
  double kernel_ms;      /* K1 render_paths, CUDA events on the launch stream          */
  double reduce_ms;      /* NCCL reduce (uecrt_render with num_gpus>1), else 0         */
  double resolve_ms;     /* K3 resolve                                                 */
  double h2d_ms, d2h_ms; /* host entry points only                                     */
  int32_t sm_count, sm_clock_khz;
  int32_t k1_blocks, k1_threads, k1_smem_bytes, k1_regs;
  int32_t tiles, tile_spheres; /* 1 tile = scene resident in shared memory */
  int32_t scan_mode;           /* UECRT_SCAN_FP32, _TENSOR or _GRID: what K1 ran with */
  int32_t reserved_;
} uecrt_stats;

/* ---- lifetime ---------------------------------------------------------------------- */
const char* uecrt_last_error(void);
int uecrt_version(void);
/* number of usable sm_100 devices (0 when none; never fails) */
int uecrt_device_count(void);

/* ---- host-buffer entry points (what the re-hosted yk::render() calls) ---------------
 * uecrt_render replaces yk::render<T>() (source.cpp:98-178): pixel x spp loop, ray_color
 * (raytracer.hpp:19-37), hittable_list::hit_impl/scatter (hittable_list.hpp:32-73),
 * sphere::hit_impl (sphere.hpp:25-48), lambertian/metal scatter (material.hpp:50-75) and
 * to_color3b (source.cpp:73-83).  rgb_out: W*H*3 bytes, top row first (source.cpp:170-171).
 * accum_out_or_null: W*H*4 floats, linear radiance SUM over the rendered samples in
 * .xyz and the sample count in .w.  With num_gpus > 1 the samples are split over the
 * devices of this process and summed by one NCCL reduce. */
int uecrt_render(const uecrt_scene* scene, const uecrt_camera* cam, const uecrt_params* params,
                 uint8_t* rgb_out, float* accum_out_or_null, uecrt_stats* stats_or_null);

/* Frees the per-device workspaces (stream, events, buffers) uecrt_render caches between calls. */
void uecrt_release(void);

/* Shared-memory tiling of the scene.  0 (default) = automatic: the whole packed scene (16 B per
 * sphere) stays resident in each CTA's shared memory while three CTAs still fit on an SM (up to
 * ~4000 spheres), otherwise it is swept in tiles of 2048 spheres through a two-buffer TMA ring.  A non-zero value forces tiles of that many
 * spheres for every scene loaded afterwards (tests use tiny tiles to exercise the ring). */
void uecrt_set_tile_spheres(uint32_t spheres_per_tile);
/* Paths per K1 thread that share one sweep of the scene; applies to scenes loaded afterwards.
 * 2 is honoured only by builds with -DUECRT_ENABLE_RAYS2 (measured slower on B200; see kernels.cu). */
void uecrt_set_rays_per_thread(int rays);

/* Process default for calls that force neither scan.  UECRT_SCAN_AUTO picks the tensor-core scan for scenes of
 * 96 spheres and more, and the FP32 scan for smaller scenes (faster there) and when tiles were forced with
 * uecrt_set_tile_spheres.  Up to uecrt_tensor_scan_resident_spheres() spheres the tensor scan keeps its sphere
 * operand (128 B per sphere) resident in shared memory; above that it sweeps it through shared memory in
 * 512-sphere slices, up to uecrt_tensor_scan_max_spheres() (the scene size limit, 2^22).  The FP32 scan keeps 16-bit
 * candidate lists and is refused above 65 534 spheres; the automatic mode uses the tensor-core scan there. */
void uecrt_set_scan_mode(int uecrt_scan_mode);
/* Which pipeline K1 runs the tensor-core scan with when the sphere operand is resident in shared memory:
 *   UECRT_TENSOR_PIPELINE_SELF      (default) every shade warp reads the TMEM accumulators of its own rays
 *                                   (csrc/tc_self.cuh: 6 groups of 128 paths, 3 accumulator lanes, 3 issuer warps);
 *   UECRT_TENSOR_PIPELINE_SCANWARPS dedicated scan warps hand candidate bitmaps to the shade warps through shared
 *                                   memory (csrc/tc_scan.cuh; always used for swept operands and by uecrt_primary_hits).
 * Both trace exactly the same rays.  The environment variable UECRT_TC_PIPELINE=self|scanwarps sets the process
 * default; this call overrides it. */
enum { UECRT_TENSOR_PIPELINE_SCANWARPS = 0, UECRT_TENSOR_PIPELINE_SELF = 1 };
void uecrt_set_tensor_pipeline(int uecrt_tensor_pipeline);
int uecrt_tensor_scan_max_spheres(void);
int uecrt_tensor_scan_resident_spheres(void);

/* uecrt_primary_hits replaces hittable_list::hit_impl (hittable_list.hpp:32-58) on caller
 * rays: closest hit in the CLOSED interval [t_min, t_max], later index wins ties.
 * o_xyz/d_xyz: n_rays*3 floats (direction need not be unit; t is in units of |d| as in
 * sphere.hpp:35).  Outputs: t (double, or -1 on miss), n_xyz (3 doubles), id (-1 on miss),
 * front_face.  exact != 0 uses the FP64-resolved path (reference-exact index); exact == 0
 * is the fast path K1 uses by default: FP32 comparison of the candidates, the winner's root and
 * normal in FP64 for the caller's direction (tolerance: see UECRT_FLAG_EXACT_HITS). */
int uecrt_primary_hits(const uecrt_scene* scene, uint64_t n_rays, const float* o_xyz,
                       const float* d_xyz, double t_min, double t_max, int exact, double* t,
                       double* n_xyz, int32_t* id, uint8_t* front_face);

/* uecrt_scatter replaces lambertian::scatter / metal::scatter (material.hpp:50-75) and the
 * RTIOW dielectric, with the random numbers supplied by the caller (uniforms in [0,1), 4
 * per call) so the result can be compared with the oracle fed the same numbers.
 * Per call i: mat_type[i], mat_param[i], incoming direction, hit point, unit normal,
 * front_face -> scattered (0/1), attenuation rgb, scattered ray origin+direction. */
int uecrt_scatter(uint64_t n, const uint32_t* mat_type, const uecrt_float4* mat_param,
                  const float* d_in_xyz, const float* p_xyz, const float* normal_xyz,
                  const uint8_t* front_face, const float* uniforms4, uint8_t* scattered,
                  float* attenuation_rgb, float* o_out_xyz, float* d_out_xyz);

/* Test hook for the tensor-core candidate scan's SUPERSET guarantee: runs the scan alone on caller rays and reports its
 * verdict before any resolve.  flags_out[i * n + j] != 0 iff the scan flagged sphere j for ray i; the scan flags whole
 * groups of *spheres_per_flag_out consecutive spheres (1 or 4), so j's flag is its group's.  Every sphere whose reference
 * discriminant (sphere.hpp:27-32, in double on the same float inputs) is >= 0 must be flagged. */
int uecrt_debug_tensor_candidates(const uecrt_scene* scene, uint64_t n_rays, const float* o_xyz,
                                  const float* d_xyz, uint8_t* flags_out, int32_t* spheres_per_flag_out);

/* Test hook for K1's camera rays (source.cpp:162-165 + camera.hpp:29-32 on the device): the ray K1 starts path
 * (pixel, sample) of a width x height frame with.  The jitter and the lens sample come from the Philox block
 * (pixel, sample, 0, 0) of `seed`: u = (x + U0) / width, v = (height - 1 - y + U1) / height with pixel = y * width + x,
 * lens point = lens_radius * sqrt(U2) * (cos 2 pi U3, sin 2 pi U3) in the camera's (u, v) basis.  o_xyz: origin (+ lens
 * offset); d_xyz: lower_left + u horizontal + v vertical - origin - offset, NOT normalised (K1 normalises it). */
int uecrt_debug_camera_rays(const uecrt_camera* cam, uint32_t width, uint32_t height, uint64_t seed, uint64_t n,
                            const uint32_t* pixel, const uint32_t* sample, float* o_xyz, float* d_xyz);

/* Philox4x32-10 block for (key = seed, counter = {c0,c1,c2,c3}), computed on the device. */
int uecrt_philox(uint64_t n, uint64_t seed, const uint32_t* counters4, uint32_t* out4);

/* ---- device-resident entry points (one process per GPU; bench.py / torch.distributed) --
 * A uecrt_ctx owns the device copy of one scene (+camera) on the current CUDA device. */
typedef struct uecrt_ctx uecrt_ctx;
int uecrt_ctx_create(const uecrt_scene* scene, const uecrt_camera* cam, uecrt_ctx** out);
void uecrt_ctx_destroy(uecrt_ctx* ctx);
/* Re-upload a scene (+camera) into an existing context: host->device copies only, device buffers
 * are reused when they fit (the per-frame H2D of a caller whose scene changes every frame). */
int uecrt_ctx_set_scene(uecrt_ctx* ctx, const uecrt_scene* scene, const uecrt_camera* cam);
/* K1: accumulate samples [sample_begin, +sample_count) of every pixel into d_accum
 * (device, W*H float4, must be zeroed by the caller before the first call of a frame). */
int uecrt_accumulate_device(uecrt_ctx* ctx, const uecrt_params* params, float* d_accum,
                            void* cuda_stream, uecrt_stats* stats_or_null);
/* K3: to_color3b (source.cpp:73-83): d_accum (sum, any .w) / spp -> sqrt -> clamp -> *256 */
int uecrt_resolve_device(const float* d_accum, uint32_t width, uint32_t height, uint32_t spp,
                         uint8_t* d_rgb, void* cuda_stream);
/* uecrt_primary_hits against the context's resident scene: no scene upload, no allocation once the context's scratch
 * has grown to the batch size.  What yk::hittable_list::hit (host/yk/hittable_list.hpp) calls, ray by ray or batched. */
int uecrt_ctx_primary_hits(uecrt_ctx* ctx, uint64_t n_rays, const float* o_xyz, const float* d_xyz,
                           double t_min, double t_max, int exact, double* t, double* n_xyz, int32_t* id,
                           uint8_t* front_face);
/* ---- frame queue: overlap what FOLLOWS the path (source.cpp:221-230: to_color3b, then the PNG write) with the next
 * frame's path.  uecrt_queue_submit is asynchronous: memset + K1 + K3 go to the queue's compute stream, the RGB image
 * is copied into pinned host memory by a copy engine on a second stream.  uecrt_queue_next blocks until the OLDEST
 * submitted frame has landed and returns a pointer to it, valid until the slot is submitted to again (two slots: at
 * most two frames in flight).  Typical loop: submit(0); for k: submit(k + 1); next(k) -> encode PNG of frame k on a host
 * thread while K1 of frame k + 1 runs.  cam_or_null changes the camera for this and later frames (no copy: the camera
 * travels in K1's argument block); the scene is the one given at creation.  One device, the caller's current one. */
typedef struct uecrt_queue uecrt_queue;
int uecrt_queue_create(const uecrt_scene* scene, const uecrt_camera* cam, uint32_t width, uint32_t height, uecrt_queue** out);
void uecrt_queue_destroy(uecrt_queue* q);
int uecrt_queue_submit(uecrt_queue* q, const uecrt_camera* cam_or_null, const uecrt_params* params, uint64_t frame_id);
int uecrt_queue_next(uecrt_queue* q, const uint8_t** rgb, uint64_t* frame_id_or_null, float* k1_ms_or_null);

/* statistics of the last uecrt_accumulate_device on ctx (blocks on the stream's events) */
int uecrt_ctx_stats(uecrt_ctx* ctx, uecrt_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* UECRT_H */
