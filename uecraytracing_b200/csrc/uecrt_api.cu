// uecrt_api.cu -- implementation of the C ABI declared in include/uecrt.h.
//
// Host side of the boundary: packs the flattened hittable_list (scene_pack.hpp), owns device
// memory and CUDA events, launches K1/K2/K3 (kernels.cu) and, for uecrt_render with
// num_gpus > 1, shards samples-per-pixel over the devices of this process (one host thread per
// GPU) and sums the float4 accumulation buffers with ONE ncclReduce over NVLink.
// There is no CPU fallback: without an sm_100 device every compute entry point fails with
// UECRT_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/uecrt.h"
#include "kernels.cuh"
#include "grid_build.hpp"
#include "scene_pack.hpp"

#ifndef UECRT_DEFAULT_RAYS
#define UECRT_DEFAULT_RAYS 1
#endif

namespace {

thread_local std::string g_err;
thread_local unsigned long long g_err_seq = 0;
std::mutex g_err_mutex;
std::string g_last_err_any;   // last error from any thread (worker threads of uecrt_render)
unsigned long long g_last_err_any_seq = 0, g_err_counter = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  {
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_last_err_any = buf;
    g_err_seq = g_last_err_any_seq = ++g_err_counter;
  }
  return code;
}

#define CUDA_TRY(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail(UECRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

bool device_is_sm100(int dev) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
  return major == 10;
}

int usable_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d)
    if (device_is_sm100(d)) ++ok; else break;
  return ok;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(UECRT_ERR_NO_DEVICE, "no CUDA device (%s); libuecrt has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (!device_is_sm100(dev)) return fail(UECRT_ERR_NO_DEVICE, "device %d is not sm_100 (B200); kernels are built for sm_100a only", dev);
  return UECRT_OK;
}

// The FP32 scan's per-thread candidate lists hold 16-bit sphere indices: scenes above kFp32ScanMaxSpheres run on the
// tensor-core scan (candidate bitmaps, swept operand) or the uniform grid (32-bit items), which have no such limit.
constexpr uint32_t kFp32ScanMaxSpheres = 65534u, kMaxSceneSpheres = 1u << 22;

int validate_scene(const uecrt_scene* s) {
  if (!s) return fail(UECRT_ERR_INVALID, "scene is null");
  if (s->n > kMaxSceneSpheres) return fail(UECRT_ERR_INVALID, "scene has %u spheres (max %u)", s->n, kMaxSceneSpheres);
  if (s->n && (!s->geom || !s->mat_type || !s->mat_param)) return fail(UECRT_ERR_INVALID, "scene arrays are null");
  for (uint32_t i = 0; i < s->n; ++i) {
    if (s->mat_type[i] > UECRT_MAT_DIELECTRIC) return fail(UECRT_ERR_INVALID, "sphere %u: unknown material type %u", i, s->mat_type[i]);
    // the scans square the coordinates (|c|^2 - r^2): beyond ~1e18 that overflows FP32 and the pre-filters lose their guarantee
    const uecrt_float4 g = s->geom[i];
    const bool ok = std::fabs(g.x) <= 1e18f && std::fabs(g.y) <= 1e18f && std::fabs(g.z) <= 1e18f && std::fabs(g.w) <= 1e18f;   // false for NaN
    if (!ok) return fail(UECRT_ERR_INVALID, "sphere %u: centre / radius must be finite and below 1e18 in magnitude", i);
  }
  return UECRT_OK;
}

}  // namespace

struct uecrt_ctx {
  int device = 0;
  uecrt::DeviceScene sc{};
  uecrt::DeviceCamera cam{};
  void* d_pairs = nullptr;
  void* d_geom = nullptr;
  void* d_mat_type = nullptr;
  void* d_mat_param = nullptr;
  unsigned long long* d_counters = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  int sm_count = 0, sm_clock_khz = 0, regs = 0;
  int blocks_per_sm[2] = {0, 0};   // [0] fast K1, [1] exact K1
  int last_blocks = 0;
  size_t smem = 0;
  uint32_t n = 0;
  size_t cap_pairs = 0;    // bytes allocated for d_pairs
  uint32_t cap_n = 0;      // spheres allocated for geom / materials
  bool configured = false;
  int rays = 1;            // paths per K1 thread
  // tensor-core scan (tc_scan.cuh): available when the scene fits its residency limit
  void* d_tc = nullptr;
  size_t cap_tc = 0;
  void* d_quads = nullptr;
  size_t cap_quads = 0;
  // grow-only scratch of uecrt_ctx_primary_hits (rays in, hit records out)
  void* d_hits = nullptr;
  size_t cap_hits = 0;
  size_t smem_tc = 0;
  int regs_tc = 0;
  // uniform grid (grid_build.hpp): built and uploaded on first use of UECRT_SCAN_GRID with the current scene
  std::vector<uecrt_float4> h_geom;
  void *d_cell_start[2] = {nullptr, nullptr}, *d_items[2] = {nullptr, nullptr}, *d_always = nullptr;   // [0] fine, [1] coarse level
  size_t cap_cells[2] = {0, 0}, cap_items[2] = {0, 0}, cap_always = 0;
  bool grid_ready = false;
  int regs_grid = 0, blocks_grid[2] = {0, 0};
  int last_scan_mode = UECRT_SCAN_FP32, last_threads = uecrt::kRenderThreads, last_regs = 0;
  size_t last_smem = 0;
};

namespace {

std::atomic<uint32_t> g_tile_spheres{0};   // 0 = automatic (uecrt_set_tile_spheres)
std::atomic<int> g_scan_mode{UECRT_SCAN_AUTO};
#ifndef UECRT_AUTO_SCAN
#define UECRT_AUTO_SCAN UECRT_SCAN_TENSOR   // what UECRT_SCAN_AUTO resolves to when the scene fits the tensor scan
#endif
std::atomic<int> g_rays_per_thread{UECRT_DEFAULT_RAYS};   // paths per K1 thread sharing one scene scan (1 or 2)
constexpr int kCounterWords = 2048;      // [0,64): statistics and profile sums; [128, 128 + 5*128): UECRT_TC_TRACE timeline
#ifndef UECRT_AUTO_SWEEP
#define UECRT_AUTO_SWEEP 1
#endif
constexpr bool kAutoSweep = UECRT_AUTO_SWEEP != 0;   // automatic mode: scenes above the residency limit use the swept tensor scan
constexpr uint32_t kTensorScanMinSpheres = 96;   // automatic mode: smaller scenes use the FP32 scan
constexpr uint32_t kAutoTileSpheres = 2048;   // 32 KB tiles: 2 buffers + lists = 73 KB -> 3 CTAs per SM
constexpr size_t kSmemOptin = 232448;

int ctx_upload_scene(uecrt_ctx* c, const uecrt_scene* scene) {
  // A scene reload overwrites (or reallocates) buffers the context's last K1 may still be reading: wait for it.
  if (c->timed && c->ev1) CUDA_TRY(cudaEventSynchronize(c->ev1));
  // resident while 3 CTAs of K1 still fit on an SM (scenes up to ~4000 spheres); beyond that the
  // scene is swept in 32 KB tiles, which keeps 3 CTAs/SM (measured on the 10k-sphere scene: 66% of
  // FP32 peak tiled vs 62% resident at 1 CTA/SM) and has no upper bound but the 16-bit list index
  uint32_t tile = g_tile_spheres.load();
  if (tile == 0) {
    const uecrt::PackedScene whole = uecrt::pack_scene(*scene, 0);
    if (uecrt::render_smem_bytes((int)whole.tile_pairs, 1, uecrt::kRenderThreads, g_rays_per_thread.load()) > kSmemOptin / 3 - 1024) tile = kAutoTileSpheres;
  }
  const uecrt::PackedScene ps = uecrt::pack_scene(*scene, tile);
  const size_t pair_bytes = ps.bytes();
  c->rays = g_rays_per_thread.load();
  const size_t smem_need = uecrt::render_smem_bytes((int)ps.tile_pairs, (int)ps.ntiles, uecrt::kRenderThreads, c->rays);
  if (smem_need > kSmemOptin)
    return fail(UECRT_ERR_INVALID, "tile of %u sphere pairs needs %zu B of shared memory per CTA (max %zu)", ps.tile_pairs, smem_need, kSmemOptin);
  const uint32_t n_alloc = std::max<uint32_t>(scene->n, 1);
  if (pair_bytes > c->cap_pairs) {             // grow-only: a reused context re-uploads without cudaMalloc
    cudaFree(c->d_pairs); c->d_pairs = nullptr; c->cap_pairs = 0;
    CUDA_TRY(cudaMalloc(&c->d_pairs, pair_bytes));
    c->cap_pairs = pair_bytes;
  }
  if (n_alloc > c->cap_n) {
    cudaFree(c->d_geom); cudaFree(c->d_mat_type); cudaFree(c->d_mat_param);
    c->d_geom = c->d_mat_type = c->d_mat_param = nullptr; c->cap_n = 0;
    CUDA_TRY(cudaMalloc(&c->d_geom, (size_t)n_alloc * sizeof(uecrt_float4)));
    CUDA_TRY(cudaMalloc(&c->d_mat_type, (size_t)n_alloc * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&c->d_mat_param, (size_t)n_alloc * sizeof(uecrt_float4)));
    c->cap_n = n_alloc;
  }
  CUDA_TRY(cudaMemcpy(c->d_pairs, ps.pairs.data(), pair_bytes, cudaMemcpyHostToDevice));
  if (scene->n) {
    CUDA_TRY(cudaMemcpy(c->d_geom, scene->geom, (size_t)scene->n * sizeof(uecrt_float4), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_mat_type, scene->mat_type, (size_t)scene->n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_mat_param, scene->mat_param, (size_t)scene->n * sizeof(uecrt_float4), cudaMemcpyHostToDevice));
  }
  // tensor-core scan operand (128 B per sphere): resident in shared memory when the scene fits, else padded to
  // whole slices of tc_sweep_rows(n) rows that the kernel sweeps through two shared-memory buffers
  c->sc.tc_rows = nullptr; c->sc.tc_quads = nullptr; c->sc.tc_nrows = 0; c->sc.tc_npasses = 1; c->smem_tc = 0;
  {
    const bool resident = (int)scene->n <= uecrt::tc_resident_max_spheres();
    const int slice_rows = resident ? uecrt::tc_chunk() : uecrt::tc_sweep_rows((int)scene->n);
    const uecrt::PackedSceneTc pt = uecrt::pack_scene_tc(*scene, (uint32_t)slice_rows);
    if (pt.bytes.size() > c->cap_tc) {
      cudaFree(c->d_tc); c->d_tc = nullptr; c->cap_tc = 0;
      CUDA_TRY(cudaMalloc(&c->d_tc, pt.bytes.size()));
      c->cap_tc = pt.bytes.size();
    }
    CUDA_TRY(cudaMemcpy(c->d_tc, pt.bytes.data(), pt.bytes.size(), cudaMemcpyHostToDevice));
    const std::vector<float> quads = uecrt::pack_scene_quads(*scene, pt.rows);
    if (quads.size() * sizeof(float) > c->cap_quads) {
      cudaFree(c->d_quads); c->d_quads = nullptr; c->cap_quads = 0;
      CUDA_TRY(cudaMalloc(&c->d_quads, quads.size() * sizeof(float)));
      c->cap_quads = quads.size() * sizeof(float);
    }
    CUDA_TRY(cudaMemcpy(c->d_quads, quads.data(), quads.size() * sizeof(float), cudaMemcpyHostToDevice));
    c->sc.tc_quads = static_cast<const float4*>(c->d_quads);
    c->sc.tc_rows = static_cast<const unsigned char*>(c->d_tc);
    c->sc.tc_nrows = (int)pt.rows;
    c->sc.tc_npasses = resident ? 1 : (int)pt.rows / slice_rows;
    c->smem_tc = uecrt::tc_smem_bytes((int)pt.rows, c->sc.tc_npasses);
  }
  // The copies above are cudaMemcpy from pageable memory on the legacy stream: they may return once the data is staged,
  // before the DMA has landed, and the launch streams (cudaStreamNonBlocking, or the caller's) do not synchronise with the
  // legacy stream.  Drain it so that every later launch, on whatever stream, sees the whole scene.
  CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
  c->h_geom.assign(scene->geom, scene->geom + scene->n);
  c->grid_ready = false;
  c->sc.grid = uecrt::DeviceGrid{};
  c->n = scene->n;
  c->sc.pairs = static_cast<const ulonglong2*>(c->d_pairs);
  c->sc.npairs = (int)ps.tile_pairs;
  c->sc.ntiles = (int)ps.ntiles;
  c->sc.n = (int)scene->n;
  c->sc.geom = static_cast<const float4*>(c->d_geom);
  c->sc.mat_type = static_cast<const uint32_t*>(c->d_mat_type);
  c->sc.mat_param = static_cast<const float4*>(c->d_mat_param);
  c->smem = smem_need;
  return UECRT_OK;
}

void ctx_free(uecrt_ctx* c) {
  if (!c) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaFree(c->d_pairs); cudaFree(c->d_geom); cudaFree(c->d_mat_type); cudaFree(c->d_mat_param); cudaFree(c->d_counters); cudaFree(c->d_tc); cudaFree(c->d_quads); cudaFree(c->d_hits);
  for (int l = 0; l < 2; ++l) { cudaFree(c->d_cell_start[l]); cudaFree(c->d_items[l]); }
  cudaFree(c->d_always);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  cudaSetDevice(prev);
  delete c;
}

// (re)load a scene + camera into a context on the current device; buffers are reused when they fit
// the camera travels by value in K1's argument block: changing it costs no copy
void ctx_set_camera(uecrt_ctx* c, const uecrt_camera* cam) {
  for (int i = 0; i < 3; ++i) {
    c->cam.origin[i] = cam->origin[i]; c->cam.ll[i] = cam->lower_left[i]; c->cam.hor[i] = cam->horizontal[i];
    c->cam.ver[i] = cam->vertical[i]; c->cam.u[i] = cam->u[i]; c->cam.v[i] = cam->v[i];
  }
  c->cam.lens = cam->lens_radius;
  c->cam.origin_is_zero = (cam->origin[0] == 0.f && cam->origin[1] == 0.f && cam->origin[2] == 0.f) ? 1 : 0;
}

int ctx_set_scene(uecrt_ctx* c, const uecrt_scene* scene, const uecrt_camera* cam) {
  int rc = validate_scene(scene);
  if (rc) return rc;
  if ((rc = ctx_upload_scene(c, scene))) return rc;
  if (cam) ctx_set_camera(c, cam);
  cudaError_t e;
  if (!c->configured) {
    if ((e = cudaMalloc(&c->d_counters, kCounterWords * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemset(c->d_counters, 0, kCounterWords * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess)
      return fail(UECRT_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
    cudaDeviceGetAttribute(&c->sm_clock_khz, cudaDevAttrClockRate, c->device);
    c->configured = true;
  }
  if ((e = uecrt::configure_kernels(&c->regs, c->blocks_per_sm, c->smem, c->sc.ntiles > 1, c->rays)) != cudaSuccess)
    return fail(UECRT_ERR_CUDA, "kernel configuration failed: %s", cudaGetErrorString(e));
  if (c->blocks_per_sm[0] < 1 || c->blocks_per_sm[1] < 1) return fail(UECRT_ERR_CUDA, "K1 does not fit on an SM (smem %zu B)", c->smem);
  if (c->sc.tc_rows && (e = uecrt::configure_kernels_tc(&c->regs_tc, c->smem_tc)) != cudaSuccess)
    return fail(UECRT_ERR_CUDA, "tensor-scan kernel configuration failed: %s", cudaGetErrorString(e));
  c->timed = false;
  return UECRT_OK;
}

int ctx_create_on_current_device(const uecrt_scene* scene, const uecrt_camera* cam, uecrt_ctx** out) {
  int rc = require_device();
  if (rc) return rc;
  uecrt_ctx* c = new uecrt_ctx();
  cudaGetDevice(&c->device);
  rc = ctx_set_scene(c, scene, cam);
  if (rc) { ctx_free(c); return rc; }
  *out = c;
  return UECRT_OK;
}

// ---- per-device workspace of the host-buffer entry point uecrt_render: stream, events, the
// accumulation / RGB buffers and a scene context are created once per process and reused, so a
// frame costs the scene H2D, K1 (+ reduce), K3 and the D2H -- no cudaMalloc/cudaFree per call.
struct Workspace {
  int device = -1;
  uecrt_ctx* ctx = nullptr;
  cudaStream_t st = nullptr;
  float* accum = nullptr; size_t accum_cap = 0;
  uint8_t* rgb = nullptr; size_t rgb_cap = 0;
  cudaEvent_t r0 = nullptr, r1 = nullptr, e0 = nullptr, e1 = nullptr, e2 = nullptr;
};
std::mutex g_ws_mutex;
std::vector<Workspace*> g_ws;

int get_workspace(int dev, size_t accum_bytes, size_t rgb_bytes, Workspace** out) {
  Workspace* w = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    for (Workspace* x : g_ws) if (x->device == dev) w = x;
    if (!w) { w = new Workspace(); w->device = dev; g_ws.push_back(w); }
  }
  CUDA_TRY(cudaSetDevice(dev));
  if (!w->st) {
    CUDA_TRY(cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&w->r0)); CUDA_TRY(cudaEventCreate(&w->r1));
    CUDA_TRY(cudaEventCreate(&w->e0)); CUDA_TRY(cudaEventCreate(&w->e1)); CUDA_TRY(cudaEventCreate(&w->e2));
  }
  if (!w->ctx) { w->ctx = new uecrt_ctx(); w->ctx->device = dev; }
  if (accum_bytes > w->accum_cap) {
    cudaFree(w->accum); w->accum = nullptr; w->accum_cap = 0;
    CUDA_TRY(cudaMalloc(&w->accum, accum_bytes)); w->accum_cap = accum_bytes;
  }
  if (rgb_bytes > w->rgb_cap) {
    cudaFree(w->rgb); w->rgb = nullptr; w->rgb_cap = 0;
    CUDA_TRY(cudaMalloc(&w->rgb, rgb_bytes)); w->rgb_cap = rgb_bytes;
  }
  *out = w;
  return UECRT_OK;
}

void release_workspaces() {
  std::lock_guard<std::mutex> lk(g_ws_mutex);
  int prev = 0;
  cudaGetDevice(&prev);
  for (Workspace* w : g_ws) {
    cudaSetDevice(w->device);
    if (w->st) { cudaStreamSynchronize(w->st); cudaStreamDestroy(w->st); }
    cudaFree(w->accum); cudaFree(w->rgb);
    for (cudaEvent_t e : {w->r0, w->r1, w->e0, w->e1, w->e2}) if (e) cudaEventDestroy(e);
    if (w->ctx) ctx_free(w->ctx);
    delete w;
  }
  g_ws.clear();
  cudaSetDevice(prev);
}

// build + upload the uniform grid of the context's current scene (first grid call after a scene load)
int ctx_ensure_grid(uecrt_ctx* c) {
  if (c->grid_ready) return UECRT_OK;
  uecrt_scene sc{c->n, c->h_geom.data(), nullptr, nullptr};
  const uecrt::HostGrids G = uecrt::build_grids(sc);
  auto up = [&](void*& dptr, size_t& cap, const std::vector<uint32_t>& v) -> int {
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(uint32_t);
    if (bytes > cap) {
      cudaFree(dptr); dptr = nullptr; cap = 0;
      CUDA_TRY(cudaMalloc(&dptr, bytes));
      cap = bytes;
    }
    if (!v.empty()) CUDA_TRY(cudaMemcpy(dptr, v.data(), v.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return UECRT_OK;
  };
  int rc;
  const uecrt::HostGrid* level[2] = {&G.fine, &G.coarse};
  for (int l = 0; l < 2; ++l)
    if ((rc = up(c->d_cell_start[l], c->cap_cells[l], level[l]->cell_start)) || (rc = up(c->d_items[l], c->cap_items[l], level[l]->items))) return rc;
  if ((rc = up(c->d_always, c->cap_always, G.always))) return rc;
  CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));      // as in ctx_upload_scene: the grid is complete before any launch reads it
  uecrt::DeviceGrid& d = c->sc.grid;
  for (int l = 0; l < 2; ++l) {
    const uecrt::HostGrid& g = *level[l];
    uecrt::DeviceGridLevel& dl = d.lvl[l];
    for (int a = 0; a < 3; ++a) { dl.org[a] = g.org[a]; dl.cell[a] = g.cell[a]; dl.inv_cell[a] = g.inv_cell[a]; dl.dim[a] = g.dim[a]; }
    dl.cell_start = g.items.empty() ? nullptr : static_cast<const uint32_t*>(c->d_cell_start[l]);
    dl.items = static_cast<const uint32_t*>(c->d_items[l]);
  }
  d.always = static_cast<const uint32_t*>(c->d_always);
  d.n_always = (int)G.always.size();
  cudaError_t e = uecrt::configure_kernels_grid(&c->regs_grid, c->blocks_grid);
  if (e != cudaSuccess || c->blocks_grid[0] < 1 || c->blocks_grid[1] < 1)
    return fail(UECRT_ERR_CUDA, "grid kernel configuration failed: %s", cudaGetErrorString(e));
  c->grid_ready = true;
  return UECRT_OK;
}

// which scan a call runs with: forced by flag, else the process default, else UECRT_AUTO_SCAN when the scene fits
int rc_mode(uint32_t flags, const uecrt_ctx* ctx, int* out) {
  int want = (flags & UECRT_FLAG_SCAN_GRID) ? UECRT_SCAN_GRID : (flags & UECRT_FLAG_SCAN_TENSOR) ? UECRT_SCAN_TENSOR
             : (flags & UECRT_FLAG_SCAN_FP32) ? UECRT_SCAN_FP32 : g_scan_mode.load();
  const bool forced = want != UECRT_SCAN_AUTO;
  // automatic: the tensor-core scan whenever the scene fits it, unless the caller forced shared-memory
  // tiles (uecrt_set_tile_spheres), which is a request for the FP32 scan's tile ring
  // and below ~100 spheres the FP32 scan is faster (one 128-column MMA chunk per tile is the tensor scan's floor;
  // measured crossover between 64 and 128 spheres, profiles/README.md)
  if (want == UECRT_SCAN_AUTO) want = (g_tile_spheres.load() || ctx->n < kTensorScanMinSpheres || (!kAutoSweep && ctx->sc.tc_npasses > 1)) ? UECRT_SCAN_FP32 : UECRT_AUTO_SCAN;
  if (want == UECRT_SCAN_FP32 && ctx->n > kFp32ScanMaxSpheres) {
    if (forced) return fail(UECRT_ERR_INVALID, "the FP32 scan holds 16-bit candidate indices (max %u spheres); this scene has %u: use the tensor-core scan or the grid", kFp32ScanMaxSpheres, ctx->n);
    want = UECRT_SCAN_TENSOR;
  }
  if (want == UECRT_SCAN_TENSOR && !ctx->sc.tc_rows) {
    if (forced) return fail(UECRT_ERR_INVALID, "tensor-core scan requested but no operand was built for this scene (%u spheres)", ctx->n);
    want = UECRT_SCAN_FP32;
  }
  *out = want;
  return UECRT_OK;
}

int accumulate(uecrt_ctx* ctx, const uecrt_params* p, float* d_accum, cudaStream_t st) {
  if (!ctx || !p || !d_accum) return fail(UECRT_ERR_INVALID, "null argument");
  if (p->width == 0 || p->height == 0) return fail(UECRT_ERR_INVALID, "empty image");
  uint32_t begin = p->sample_begin, count = p->sample_count;
  if (count == 0) { begin = 0; count = p->spp; }
  uecrt::RenderArgs a{};
  a.sc = ctx->sc; a.cam = ctx->cam;
  a.width = p->width; a.height = p->height;
  a.sample_begin = begin; a.sample_count = count;
  a.max_depth = p->max_depth;
  a.key0 = (uint32_t)(p->seed & 0xffffffffull); a.key1 = (uint32_t)(p->seed >> 32);
  a.tiles_x = (p->width + 7) / 8;
  a.inv_tiles_x = 1.0f / (float)a.tiles_x;
  a.n_slots = (unsigned long long)a.tiles_x * ((p->height + 3) / 4) * 32ull;
  a.total_items = a.n_slots * count;   // one work item per (pixel slot, sample) path
  a.item_begin = 0;
  if (p->shard_count) {                // the shard_index-th of shard_count equal ranges of those items
    if (p->shard_index >= p->shard_count) return fail(UECRT_ERR_INVALID, "shard_index %u >= shard_count %u", p->shard_index, p->shard_count);
    const unsigned long long t = a.total_items, lo = t / p->shard_count * p->shard_index + t % p->shard_count * p->shard_index / p->shard_count,
                             hi = t / p->shard_count * (p->shard_index + 1) + t % p->shard_count * (p->shard_index + 1) / p->shard_count;
    a.item_begin = lo; a.total_items = hi - lo;
  }
  a.accum = reinterpret_cast<float4*>(d_accum);
  a.counters = ctx->d_counters;
  a.count_stats = (p->flags & UECRT_FLAG_NO_COUNTERS) ? 0 : 1;
  CUDA_TRY(cudaMemsetAsync(ctx->d_counters, 0, kCounterWords * sizeof(unsigned long long), st));
  const bool exact = (p->flags & UECRT_FLAG_EXACT_HITS) != 0;
  int mode = 0;
  if ((rc_mode(p->flags, ctx, &mode))) return UECRT_ERR_INVALID;
  const bool tensor = mode == UECRT_SCAN_TENSOR, grid = mode == UECRT_SCAN_GRID;
  if (grid) {
    int rc = ctx_ensure_grid(ctx);
    if (rc) return rc;
    a.sc = ctx->sc;
  }
  const int blocks = tensor ? ctx->sm_count : ctx->sm_count * (grid ? ctx->blocks_grid[exact ? 1 : 0] : ctx->blocks_per_sm[exact ? 1 : 0]);
  ctx->last_blocks = blocks;
  ctx->last_scan_mode = mode;
  ctx->last_threads = uecrt::kRenderThreads;
  ctx->last_smem = grid ? 0 : ctx->smem;
  if (tensor) {
    int thr = 0; size_t sm = 0;
    uecrt::tc_render_config(a.sc, &thr, &sm, nullptr);
    ctx->last_threads = thr; ctx->last_smem = sm;
  }
  ctx->last_regs = tensor ? ctx->regs_tc : grid ? ctx->regs_grid : ctx->regs;
  CUDA_TRY(cudaEventRecord(ctx->ev0, st));
  if (tensor) uecrt::launch_render_paths_tc(a, exact, blocks, ctx->smem_tc, st);
  else if (grid) uecrt::launch_render_paths_grid(a, exact, blocks, st);
  else uecrt::launch_render_paths(a, exact, ctx->rays, blocks, uecrt::kRenderThreads, ctx->smem, st);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(ctx->ev1, st));
  ctx->timed = true;
  return UECRT_OK;
}

int ctx_stats(uecrt_ctx* ctx, uecrt_stats* out) {
  if (!ctx || !out) return fail(UECRT_ERR_INVALID, "null argument");
  std::memset(out, 0, sizeof *out);
  out->sm_count = ctx->sm_count; out->sm_clock_khz = ctx->sm_clock_khz;
  out->k1_blocks = ctx->last_blocks; out->k1_threads = ctx->last_threads;
  out->k1_smem_bytes = (int32_t)(ctx->last_scan_mode == UECRT_SCAN_GRID ? 0 : ctx->last_smem ? ctx->last_smem : ctx->smem);
  out->k1_regs = ctx->last_regs ? ctx->last_regs : ctx->regs;
  out->scan_mode = ctx->last_scan_mode;
  out->tiles = ctx->sc.ntiles; out->tile_spheres = ctx->sc.npairs * 2;
  if (!ctx->timed) return UECRT_OK;
  CUDA_TRY(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  out->kernel_ms = ms;
  unsigned long long h[4];
  CUDA_TRY(cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
  out->rays_traced = h[1]; out->paths = h[2];
  out->sphere_tests = ctx->last_scan_mode == UECRT_SCAN_GRID ? h[3] : h[1] * (unsigned long long)ctx->n;
#ifdef UECRT_TC_TRACE
  {   // timeline of one CTA's issuer and its four scan warps over 48 consecutive chunks (debugging builds only)
    static unsigned long long q[kCounterWords];
    CUDA_TRY(cudaMemcpy(q, ctx->d_counters, sizeof q, cudaMemcpyDeviceToHost));
    if (q[128]) {
      const unsigned long long t0 = q[128];
      fprintf(stderr, "[tc trace] CTA 0, chunk sequence numbers from 3000; cycles relative to the first event\n"
                      "[tc trace]  seq | issuer: reaches the slot wait, slot free, MMAs+commit issued | scan warp 0..3: slot full seen, slot released, packed\n");
      for (int i = 0; i < 48; ++i) {
        const unsigned long long* e = q + 128 + 20 * i;
        fprintf(stderr, "[tc trace] %4d | %6lld %6lld %6lld |", 3000 + i, (long long)(e[0] - t0), (long long)(e[1] - t0), (long long)(e[2] - t0));
        for (int w = 0; w < 4; ++w) fprintf(stderr, " %6lld %6lld %6lld |", (long long)(e[4 + 4 * w] - t0), (long long)(e[5 + 4 * w] - t0), (long long)(e[6 + 4 * w] - t0));
        fprintf(stderr, "\n");
      }
    }
  }
#endif
#ifdef UECRT_TC_PROFILE
  {   // phase clocks of the tensor-scan K1 shade warps (SM cycles per warp-round), debugging builds only
    static unsigned long long q[kCounterWords];
    CUDA_TRY(cudaMemcpy(q, ctx->d_counters, sizeof q, cudaMemcpyDeviceToHost));
    if (q[4]) {
      const char* nm[7] = {"full_wait", "prepare_next", "submit", "wait_scan", "resolve", "shade", "total"};
      fprintf(stderr, "[tc profile] warp-rounds %llu:", q[4]);
      for (int i = 0; i < 7; ++i) fprintf(stderr, " %s %.0f", nm[i], (double)q[5 + i] / (double)q[4]);
      fprintf(stderr, " cycles/round\n");
      fprintf(stderr, "[tc profile] per warp-round: candidate trips %.2f, inner word steps %.2f, candidates (lane sum) %.2f\n", (double)q[16] / q[4],
              (double)q[17] / q[4], (double)q[18] / q[4]);
      if (q[20]) fprintf(stderr, "[tc profile] issuer: tiles %llu, per tile: wait a_ready %.0f, wait tmem_empty %.0f, issue %.0f cycles\n", q[20],
                         (double)q[21] / q[20], (double)q[22] / q[20], (double)q[23] / q[20]);
      if (q[24] + q[25]) fprintf(stderr, "[tc profile] scan warps: %.1f %% of the chunk waits found the MMAs already complete\n", 100.0 * q[24] / (double)(q[24] + q[25]));
      fprintf(stderr, "[tc profile] blocked chunk waits by chunk index: %llu %llu %llu %llu (of %llu tiles x 4 warps)\n", q[26], q[27], q[28], q[29], q[20]);
      if (q[12]) fprintf(stderr, "[tc profile] scan warps: tiles %llu, per tile: wait_full %.0f ld+pack %.0f arrive %.0f cycles\n", q[12],
                         (double)q[13] / q[12], (double)q[14] / q[12], (double)q[15] / q[12]);
      if (q[40] + q[41]) fprintf(stderr, "[tc profile] issuer, swept operand, per tile: wait b_free %.0f, load + wait b_full %.0f cycles\n", (double)q[40] / q[20], (double)q[41] / q[20]);
      fprintf(stderr, "[tc profile] blocked chunk-0 waits by group: count %llu %llu %llu %llu, mean cycles %.0f %.0f %.0f %.0f\n", q[42], q[43], q[44], q[45],
              q[42] ? (double)q[46] / q[42] : 0.0, q[43] ? (double)q[47] / q[43] : 0.0, q[44] ? (double)q[48] / q[44] : 0.0, q[45] ? (double)q[49] / q[45] : 0.0);
      if (q[12]) fprintf(stderr, "[tc profile] per scan warp (SM sub-partition 0..3), per tile: wait_full %.0f %.0f %.0f %.0f, ld+pack %.0f %.0f %.0f %.0f\n",
                         4.0 * q[32] / q[12], 4.0 * q[33] / q[12], 4.0 * q[34] / q[12], 4.0 * q[35] / q[12],
                         4.0 * q[36] / q[12], 4.0 * q[37] / q[12], 4.0 * q[38] / q[12], 4.0 * q[39] / q[12]);
    }
  }
#endif
  return UECRT_OK;
}

// ---- NCCL through dlopen: only uecrt_render(num_gpus > 1) needs it ---------------------
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*Reduce)(const void* send, void* recv, size_t count, int dtype, int op, int root, void* comm, cudaStream_t st) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat = 7, kNcclSum = 0;   // ncclFloat32, ncclSum (nccl.h enums, stable since 2.0)

int load_nccl(NcclApi& n) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.lib) break;
  }
  if (!n.lib) return fail(UECRT_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(dlsym(n.lib, "ncclCommInitAll"));
  n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.lib, "ncclCommDestroy"));
  n.Reduce = reinterpret_cast<decltype(n.Reduce)>(dlsym(n.lib, "ncclReduce"));
  n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(dlsym(n.lib, "ncclGroupStart"));
  n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(dlsym(n.lib, "ncclGroupEnd"));
  n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.lib, "ncclGetErrorString"));
  if (!n.CommInitAll || !n.CommDestroy || !n.Reduce || !n.GroupStart || !n.GroupEnd)
    return fail(UECRT_ERR_NCCL, "libnccl lacks a required symbol");
  return UECRT_OK;
}

// process-global communicator cache (communicator creation is excluded from frame timing)
struct CommCache {
  std::mutex m;
  NcclApi api;
  int ndev = 0;
  std::vector<void*> comms;
} g_comms;

int get_comms(int ndev, std::vector<void*>& out, NcclApi*& api) {
  std::lock_guard<std::mutex> lk(g_comms.m);
  if (!g_comms.api.lib) {
    int rc = load_nccl(g_comms.api);
    if (rc) return rc;
  }
  if (g_comms.ndev != ndev) {
    for (void* c : g_comms.comms) g_comms.api.CommDestroy(c);
    g_comms.comms.assign(ndev, nullptr);
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; ++i) devs[i] = i;
    int r = g_comms.api.CommInitAll(g_comms.comms.data(), ndev, devs.data());
    if (r != 0) {
      g_comms.comms.clear(); g_comms.ndev = 0;
      return fail(UECRT_ERR_NCCL, "ncclCommInitAll(%d) failed: %s", ndev, g_comms.api.GetErrorString ? g_comms.api.GetErrorString(r) : "?");
    }
    g_comms.ndev = ndev;
  }
  out = g_comms.comms;
  api = &g_comms.api;
  return UECRT_OK;
}

}  // namespace

// =========================================================================================
extern "C" {

// The most recent failure of the calling thread's last API call; if that call failed inside a worker thread
// (uecrt_render with num_gpus > 1) the worker's message.  Every failing path goes through fail(), which stamps both.
const char* uecrt_last_error(void) {
  std::lock_guard<std::mutex> lk(g_err_mutex);
  if (g_err_seq < g_last_err_any_seq || g_err.empty()) { g_err = g_last_err_any; g_err_seq = g_last_err_any_seq; }
  return g_err.c_str();
}
int uecrt_version(void) { return UECRT_VERSION; }
int uecrt_device_count(void) { return usable_device_count(); }

int uecrt_ctx_create(const uecrt_scene* scene, const uecrt_camera* cam, uecrt_ctx** out) {
  if (!out) return fail(UECRT_ERR_INVALID, "out is null");
  *out = nullptr;
  return ctx_create_on_current_device(scene, cam, out);
}
void uecrt_ctx_destroy(uecrt_ctx* ctx) { ctx_free(ctx); }
int uecrt_ctx_set_scene(uecrt_ctx* ctx, const uecrt_scene* scene, const uecrt_camera* cam) {
  if (!ctx) return fail(UECRT_ERR_INVALID, "ctx is null");
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev != ctx->device) return fail(UECRT_ERR_INVALID, "context lives on device %d, current device is %d", ctx->device, dev);
  return ctx_set_scene(ctx, scene, cam);
}

int uecrt_accumulate_device(uecrt_ctx* ctx, const uecrt_params* params, float* d_accum, void* cuda_stream,
                            uecrt_stats* stats_or_null) {
  int rc = accumulate(ctx, params, d_accum, static_cast<cudaStream_t>(cuda_stream));
  if (rc) return rc;
  if (stats_or_null) return ctx_stats(ctx, stats_or_null);
  return UECRT_OK;
}

int uecrt_resolve_device(const float* d_accum, uint32_t width, uint32_t height, uint32_t spp, uint8_t* d_rgb,
                         void* cuda_stream) {
  int rc = require_device();
  if (rc) return rc;
  if (!d_accum || !d_rgb || spp == 0) return fail(UECRT_ERR_INVALID, "bad argument");
  uecrt::launch_resolve(reinterpret_cast<const float4*>(d_accum), width * height, spp, d_rgb, static_cast<cudaStream_t>(cuda_stream));
  CUDA_TRY(cudaGetLastError());
  return UECRT_OK;
}

int uecrt_ctx_stats(uecrt_ctx* ctx, uecrt_stats* out) { return ctx_stats(ctx, out); }

// ---- frame queue: what follows the path in the reference (source.cpp:221-230: to_color3b, then the PNG write) overlapped
// with the NEXT frame's path.  Two slots; per frame the compute stream runs memset, K1, K3 and records an event, the copy
// stream waits for it and moves the RGB image into PINNED host memory with cudaMemcpyAsync (a copy engine, no SM), so the
// D2H of frame k and whatever the caller does with it on the host (PNG deflate) run while K1 of frame k + 1 occupies the SMs.
struct uecrt_queue {
  uecrt_ctx* ctx = nullptr;
  int device = 0;
  uint32_t width = 0, height = 0;
  cudaStream_t compute = nullptr, copy = nullptr;
  struct Slot {
    float* accum = nullptr;
    uint8_t* d_rgb = nullptr;
    uint8_t* h_rgb = nullptr;            // pinned
    cudaEvent_t k1_begin = nullptr, k1_end = nullptr, resolved = nullptr, copied = nullptr;
    uint64_t frame_id = 0;
  } slot[2];
  uint64_t submitted = 0, taken = 0;
};

static void queue_free(uecrt_queue* q) {
  if (!q) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(q->device);
  if (q->compute) cudaStreamSynchronize(q->compute);
  if (q->copy) cudaStreamSynchronize(q->copy);
  for (auto& s : q->slot) {
    cudaFree(s.accum); cudaFree(s.d_rgb);
    if (s.h_rgb) cudaFreeHost(s.h_rgb);
    for (cudaEvent_t e : {s.k1_begin, s.k1_end, s.resolved, s.copied}) if (e) cudaEventDestroy(e);
  }
  if (q->compute) cudaStreamDestroy(q->compute);
  if (q->copy) cudaStreamDestroy(q->copy);
  if (q->ctx) ctx_free(q->ctx);
  cudaSetDevice(prev);
  delete q;
}

int uecrt_queue_create(const uecrt_scene* scene, const uecrt_camera* cam, uint32_t width, uint32_t height, uecrt_queue** out) {
  if (!out) return fail(UECRT_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!scene || !cam || width == 0 || height == 0) return fail(UECRT_ERR_INVALID, "bad argument");
  uecrt_queue* q = new uecrt_queue();
  q->width = width; q->height = height;
  int rc = ctx_create_on_current_device(scene, cam, &q->ctx);
  if (rc) { delete q; return rc; }
  q->device = q->ctx->device;
  const size_t npx = (size_t)width * height;
  auto body = [&]() -> int {
    CUDA_TRY(cudaStreamCreateWithFlags(&q->compute, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&q->copy, cudaStreamNonBlocking));
    for (auto& s : q->slot) {
      CUDA_TRY(cudaMalloc(&s.accum, npx * 4 * sizeof(float)));
      CUDA_TRY(cudaMalloc(&s.d_rgb, npx * 3));
      CUDA_TRY(cudaHostAlloc(&s.h_rgb, npx * 3, cudaHostAllocDefault));
      CUDA_TRY(cudaEventCreate(&s.k1_begin)); CUDA_TRY(cudaEventCreate(&s.k1_end));
      CUDA_TRY(cudaEventCreateWithFlags(&s.resolved, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    }
    return UECRT_OK;
  };
  if ((rc = body())) { queue_free(q); return rc; }
  *out = q;
  return UECRT_OK;
}
void uecrt_queue_destroy(uecrt_queue* q) { queue_free(q); }

int uecrt_queue_submit(uecrt_queue* q, const uecrt_camera* cam_or_null, const uecrt_params* params, uint64_t frame_id) {
  if (!q || !params) return fail(UECRT_ERR_INVALID, "null argument");
  if (params->width != q->width || params->height != q->height) return fail(UECRT_ERR_INVALID, "the queue was created for %u x %u frames", q->width, q->height);
  if (params->spp == 0) return fail(UECRT_ERR_INVALID, "spp must be > 0");
  if (q->submitted - q->taken >= 2) return fail(UECRT_ERR_INVALID, "both slots are in flight: call uecrt_queue_next first");
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev != q->device) return fail(UECRT_ERR_INVALID, "queue lives on device %d, current device is %d", q->device, dev);
  uecrt_queue::Slot& s = q->slot[q->submitted & 1u];
  if (cam_or_null) ctx_set_camera(q->ctx, cam_or_null);
  const size_t npx = (size_t)q->width * q->height;
  CUDA_TRY(cudaMemsetAsync(s.accum, 0, npx * 4 * sizeof(float), q->compute));
  CUDA_TRY(cudaEventRecord(s.k1_begin, q->compute));
  int rc = accumulate(q->ctx, params, s.accum, q->compute);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(s.k1_end, q->compute));
  uecrt::launch_resolve(reinterpret_cast<const float4*>(s.accum), (uint32_t)npx, params->spp, s.d_rgb, q->compute);   // K3: 0.01 ms
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(s.resolved, q->compute));
  CUDA_TRY(cudaStreamWaitEvent(q->copy, s.resolved, 0));
  CUDA_TRY(cudaMemcpyAsync(s.h_rgb, s.d_rgb, npx * 3, cudaMemcpyDeviceToHost, q->copy));
  CUDA_TRY(cudaEventRecord(s.copied, q->copy));
  s.frame_id = frame_id;
  ++q->submitted;
  return UECRT_OK;
}

int uecrt_queue_next(uecrt_queue* q, const uint8_t** rgb, uint64_t* frame_id_or_null, float* k1_ms_or_null) {
  if (!q || !rgb) return fail(UECRT_ERR_INVALID, "null argument");
  if (q->taken == q->submitted) return fail(UECRT_ERR_INVALID, "no frame in flight");
  uecrt_queue::Slot& s = q->slot[q->taken & 1u];
  CUDA_TRY(cudaEventSynchronize(s.copied));
  if (k1_ms_or_null) CUDA_TRY(cudaEventElapsedTime(k1_ms_or_null, s.k1_begin, s.k1_end));
  if (frame_id_or_null) *frame_id_or_null = s.frame_id;
  *rgb = s.h_rgb;
  ++q->taken;
  return UECRT_OK;
}


int uecrt_render(const uecrt_scene* scene, const uecrt_camera* cam, const uecrt_params* params, uint8_t* rgb_out,
                 float* accum_out_or_null, uecrt_stats* stats_or_null) {
  static std::mutex render_mutex;                 // one frame at a time per process
  std::lock_guard<std::mutex> render_lock(render_mutex);
  int rc = require_device();
  if (rc) return rc;
  if (!scene || !cam || !params || !rgb_out) return fail(UECRT_ERR_INVALID, "null argument");
  if (params->spp == 0) return fail(UECRT_ERR_INVALID, "spp must be > 0");
  if (params->width == 0 || params->height == 0) return fail(UECRT_ERR_INVALID, "empty image");
  if ((rc = validate_scene(scene))) return rc;
  const int avail = usable_device_count();
  int ndev = params->num_gpus <= 0 ? 1 : params->num_gpus;
  if (ndev > avail) return fail(UECRT_ERR_INVALID, "num_gpus=%d but only %d sm_100 device(s) are visible", ndev, avail);
  uint32_t begin = params->sample_begin, count = params->sample_count;
  if (count == 0) { begin = 0; count = params->spp; }
  const size_t npx = (size_t)params->width * params->height;

  int dev0 = 0;
  CUDA_TRY(cudaGetDevice(&dev0));
  std::vector<void*> comms;
  NcclApi* nccl = nullptr;
  if (ndev > 1 && (rc = get_comms(ndev, comms, nccl))) return rc;

  std::vector<Workspace*> ws(ndev, nullptr);
  std::vector<int> wrc(ndev, 0);
  // per device: scene H2D, zero the accumulation buffer, launch K1 on this device's spp shard
  auto worker = [&](int g) {
    auto body = [&]() -> int {
      const int dev = ndev > 1 ? g : dev0;
      int r = get_workspace(dev, npx * 4 * sizeof(float), g == 0 ? npx * 3 : 0, &ws[g]);
      if (r) return r;
      Workspace& w = *ws[g];
      if ((r = ctx_set_scene(w.ctx, scene, cam))) return r;
      CUDA_TRY(cudaMemsetAsync(w.accum, 0, npx * 4 * sizeof(float), w.st));
      uecrt_params p = *params;
      // shard g of ndev of the frame's work items (GLOBAL pixel / sample ids, so the image does not depend on ndev)
      p.sample_begin = begin; p.sample_count = count;
      p.shard_index = (uint32_t)g; p.shard_count = (uint32_t)ndev;
      if (params->shard_count) {       // the caller's own shard, split once more over this process's devices
        p.shard_index = params->shard_index * (uint32_t)ndev + (uint32_t)g; p.shard_count = params->shard_count * (uint32_t)ndev;
      }
      w.ctx->timed = false;
      if ((r = accumulate(w.ctx, &p, w.accum, w.st))) return r;
      return UECRT_OK;
    };
    wrc[g] = body();
  };
  if (ndev == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int g = 0; g < ndev; ++g) th.emplace_back(worker, g);
    for (auto& t : th) t.join();
  }
  // leave nothing in flight on any path out of this function: the workspaces (and their buffers) are reused by the next call
  auto drain = [&]() {
    for (int g = 0; g < ndev; ++g)
      if (ws[g] && ws[g]->st) { cudaSetDevice(ws[g]->device); cudaStreamSynchronize(ws[g]->st); }
  };
  auto restore = [&]() { cudaSetDevice(dev0); };
  for (int g = 0; g < ndev; ++g)
    if (wrc[g]) { drain(); restore(); return wrc[g]; }

  // ---- one NCCL reduce of W*H*4 floats to device 0 (in place on the root) over NVLink
  double resolve_ms = 0, d2h_ms = 0, reduce_ms = 0;
  if (ndev > 1) {
    nccl->GroupStart();
    for (int g = 0; g < ndev; ++g) {
      cudaSetDevice(g);
      cudaEventRecord(ws[g]->r0, ws[g]->st);
      int r = nccl->Reduce(ws[g]->accum, ws[g]->accum, npx * 4, kNcclFloat, kNcclSum, 0, comms[g], ws[g]->st);
      if (r != 0) rc = fail(UECRT_ERR_NCCL, "ncclReduce failed on device %d: %s", g, nccl->GetErrorString ? nccl->GetErrorString(r) : "?");
    }
    int r = nccl->GroupEnd();
    if (r != 0 && !rc) rc = fail(UECRT_ERR_NCCL, "ncclGroupEnd failed: %s", nccl->GetErrorString ? nccl->GetErrorString(r) : "?");
    for (int g = 0; g < ndev; ++g) { cudaSetDevice(g); cudaEventRecord(ws[g]->r1, ws[g]->st); }
    if (rc) { drain(); restore(); return rc; }
  }

  // ---- root: K3 resolve + copies back
  auto root = [&]() -> int {
    CUDA_TRY(cudaSetDevice(ndev > 1 ? 0 : dev0));
    Workspace& w = *ws[0];
    CUDA_TRY(cudaEventRecord(w.e0, w.st));
    uecrt::launch_resolve(reinterpret_cast<const float4*>(w.accum), (uint32_t)npx, count, w.rgb, w.st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(w.e1, w.st));
    CUDA_TRY(cudaMemcpyAsync(rgb_out, w.rgb, npx * 3, cudaMemcpyDeviceToHost, w.st));
    if (accum_out_or_null) CUDA_TRY(cudaMemcpyAsync(accum_out_or_null, w.accum, npx * 4 * sizeof(float), cudaMemcpyDeviceToHost, w.st));
    CUDA_TRY(cudaEventRecord(w.e2, w.st));
    CUDA_TRY(cudaStreamSynchronize(w.st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, w.e0, w.e1)); resolve_ms = ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, w.e1, w.e2)); d2h_ms = ms;
    if (ndev > 1) { CUDA_TRY(cudaEventElapsedTime(&ms, w.r0, w.r1)); reduce_ms = ms; }
    return UECRT_OK;
  };
  rc = root();
  if (rc) drain();
  for (int g = 1; g < ndev; ++g) { cudaSetDevice(g); cudaStreamSynchronize(ws[g]->st); }
  if (!rc && stats_or_null) {
    uecrt_stats total{};
    for (int g = 0; g < ndev && !rc; ++g) {
      cudaSetDevice(ndev > 1 ? g : dev0);
      uecrt_stats s{};
      rc = ctx_stats(ws[g]->ctx, &s);
      if (g == 0) total = s;
      else {
        total.rays_traced += s.rays_traced; total.sphere_tests += s.sphere_tests; total.paths += s.paths;
        total.kernel_ms = std::max(total.kernel_ms, s.kernel_ms);
      }
    }
    total.reduce_ms = reduce_ms; total.resolve_ms = resolve_ms; total.d2h_ms = d2h_ms;
    *stats_or_null = total;
  }
  restore();
  return rc;
}

void uecrt_release(void) { release_workspaces(); }
void uecrt_set_tile_spheres(uint32_t spheres_per_tile) { g_tile_spheres.store(spheres_per_tile); }
void uecrt_set_tensor_pipeline(int p) { uecrt::tc_set_pipeline(p); }
void uecrt_set_scan_mode(int m) { g_scan_mode.store((m == UECRT_SCAN_FP32 || m == UECRT_SCAN_TENSOR || m == UECRT_SCAN_GRID) ? m : UECRT_SCAN_AUTO); }
int uecrt_tensor_scan_max_spheres(void) { return (int)kMaxSceneSpheres; }
int uecrt_tensor_scan_resident_spheres(void) { return uecrt::tc_resident_max_spheres(); }
void uecrt_set_rays_per_thread(int rays) { g_rays_per_thread.store((rays == 2 && uecrt::max_rays_per_thread() == 2) ? 2 : 1); }

// closest hits of caller rays against the context's scene; the device scratch lives in the context (grow-only)
static int ctx_primary_hits(uecrt_ctx* c, uint64_t n_rays, const float* o_xyz, const float* d_xyz, double t_min, double t_max,
                            int exact, double* t, double* n_xyz, int32_t* id, uint8_t* front_face) {
  if (n_rays == 0) return UECRT_OK;
  if (!o_xyz || !d_xyz || !t || !n_xyz || !id || !front_face) return fail(UECRT_ERR_INVALID, "null argument");
  // one allocation: [o: 3 f32][d: 3 f32][t: f64][n: 3 f64][id: i32][ff: u8] per ray, each array 256-byte aligned
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t so = up(n_rays * 12), sd = so, st_ = up(n_rays * 8), sn = up(n_rays * 24), si = up(n_rays * 4), sf = up(n_rays);
  const size_t need = so + sd + st_ + sn + si + sf;
  if (need > c->cap_hits) {
    cudaFree(c->d_hits); c->d_hits = nullptr; c->cap_hits = 0;
    CUDA_TRY(cudaMalloc(&c->d_hits, need));
    c->cap_hits = need;
  }
  char* base = static_cast<char*>(c->d_hits);
  float* d_o = reinterpret_cast<float*>(base); float* d_d = reinterpret_cast<float*>(base + so);
  double* d_t = reinterpret_cast<double*>(base + so + sd); double* d_n = reinterpret_cast<double*>(base + so + sd + st_);
  int32_t* d_id = reinterpret_cast<int32_t*>(base + so + sd + st_ + sn); uint8_t* d_ff = reinterpret_cast<uint8_t*>(base + so + sd + st_ + sn + si);
  CUDA_TRY(cudaMemcpy(d_o, o_xyz, n_rays * 3 * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d_d, d_xyz, n_rays * 3 * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
  int mode = 0;
  if (rc_mode(0, c, &mode)) return UECRT_ERR_INVALID;
  if (mode == UECRT_SCAN_GRID) {
    int r = ctx_ensure_grid(c);
    if (r) return r;
    uecrt::launch_primary_hits_grid(c->sc, n_rays, d_o, d_d, t_min, t_max, exact != 0, d_t, d_n, d_id, d_ff, 0);
  } else if (mode == UECRT_SCAN_TENSOR) uecrt::launch_primary_hits_tc(c->sc, n_rays, d_o, d_d, t_min, t_max, exact != 0, d_t, d_n, d_id, d_ff, c->sm_count, 0);
  else uecrt::launch_primary_hits(c->sc, n_rays, d_o, d_d, t_min, t_max, exact != 0, d_t, d_n, d_id, d_ff, 0);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(t, d_t, n_rays * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(n_xyz, d_n, n_rays * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(id, d_id, n_rays * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(front_face, d_ff, n_rays, cudaMemcpyDeviceToHost));
  return UECRT_OK;
}

int uecrt_ctx_primary_hits(uecrt_ctx* ctx, uint64_t n_rays, const float* o_xyz, const float* d_xyz, double t_min,
                           double t_max, int exact, double* t, double* n_xyz, int32_t* id, uint8_t* front_face) {
  if (!ctx) return fail(UECRT_ERR_INVALID, "ctx is null");
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev != ctx->device) return fail(UECRT_ERR_INVALID, "context lives on device %d, current device is %d", ctx->device, dev);
  return ctx_primary_hits(ctx, n_rays, o_xyz, d_xyz, t_min, t_max, exact, t, n_xyz, id, front_face);
}

// Test hook: the tensor-core pre-filter's verdict on caller rays, before any resolve.  flags_out[i * n + j] = 1 iff the
// scan flagged sphere j for ray i (with one bit per quad: iff it flagged j's quad).
int uecrt_debug_camera_rays(const uecrt_camera* cam, uint32_t width, uint32_t height, uint64_t seed, uint64_t n,
                            const uint32_t* pixel, const uint32_t* sample, float* o_xyz, float* d_xyz) {
  int rc = require_device();
  if (rc) return rc;
  if (n == 0) return UECRT_OK;
  if (!cam || !pixel || !sample || !o_xyz || !d_xyz || width == 0 || height == 0) return fail(UECRT_ERR_INVALID, "bad argument");
  uecrt_ctx tmp;                                       // only its camera conversion is used
  ctx_set_camera(&tmp, cam);
  uecrt::RenderArgs a{};
  a.cam = tmp.cam; a.width = width; a.height = height; a.max_depth = 1;
  a.key0 = (uint32_t)(seed & 0xffffffffull); a.key1 = (uint32_t)(seed >> 32);
  uint32_t *d_p = nullptr, *d_s = nullptr;
  float *d_o = nullptr, *d_d = nullptr;
  auto body = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_p, n * 4)); CUDA_TRY(cudaMalloc(&d_s, n * 4));
    CUDA_TRY(cudaMalloc(&d_o, n * 12)); CUDA_TRY(cudaMalloc(&d_d, n * 12));
    CUDA_TRY(cudaMemcpy(d_p, pixel, n * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_s, sample, n * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
    uecrt::launch_camera_rays(a, n, d_p, d_s, d_o, d_d, 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(o_xyz, d_o, n * 12, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(d_xyz, d_d, n * 12, cudaMemcpyDeviceToHost));
    return UECRT_OK;
  };
  rc = body();
  cudaFree(d_p); cudaFree(d_s); cudaFree(d_o); cudaFree(d_d);
  return rc;
}

int uecrt_debug_tensor_candidates(const uecrt_scene* scene, uint64_t n_rays, const float* o_xyz, const float* d_xyz,
                                  uint8_t* flags_out, int32_t* spheres_per_flag_out) {
  int rc = require_device();
  if (rc) return rc;
  if (spheres_per_flag_out) *spheres_per_flag_out = uecrt::tc_quad();
  if (n_rays == 0) return UECRT_OK;
  if (!scene || !o_xyz || !d_xyz || !flags_out) return fail(UECRT_ERR_INVALID, "null argument");
  uecrt_ctx* c = nullptr;
  if ((rc = ctx_create_on_current_device(scene, nullptr, &c))) return rc;
  float *d_o = nullptr, *d_d = nullptr;
  double *d_t = nullptr, *d_n = nullptr;
  int32_t* d_id = nullptr;
  uint8_t* d_ff = nullptr;
  uint32_t* d_w = nullptr;
  const int bits_per_pass = c->sc.tc_nrows / c->sc.tc_npasses / uecrt::tc_quad();
  const size_t words_per_ray = (size_t)c->sc.tc_npasses * (bits_per_pass / 32);
  std::vector<uint32_t> words(n_rays * words_per_ray);
  auto body = [&]() -> int {
    if (!c->sc.tc_rows) return fail(UECRT_ERR_INVALID, "no tensor-scan operand for this scene");
    CUDA_TRY(cudaMalloc(&d_o, n_rays * 12)); CUDA_TRY(cudaMalloc(&d_d, n_rays * 12));
    CUDA_TRY(cudaMalloc(&d_t, n_rays * 8)); CUDA_TRY(cudaMalloc(&d_n, n_rays * 24));
    CUDA_TRY(cudaMalloc(&d_id, n_rays * 4)); CUDA_TRY(cudaMalloc(&d_ff, n_rays));
    CUDA_TRY(cudaMalloc(&d_w, words.size() * 4));
    CUDA_TRY(cudaMemset(d_w, 0, words.size() * 4));
    CUDA_TRY(cudaMemcpy(d_o, o_xyz, n_rays * 12, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_d, d_xyz, n_rays * 12, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
    uecrt::launch_primary_hits_tc(c->sc, n_rays, d_o, d_d, 0.001, 1e300, false, d_t, d_n, d_id, d_ff, c->sm_count, 0, d_w);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(words.data(), d_w, words.size() * 4, cudaMemcpyDeviceToHost));
    return UECRT_OK;
  };
  rc = body();
  cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_n); cudaFree(d_id); cudaFree(d_ff); cudaFree(d_w);
  if (!rc) {
    const uint32_t n = scene->n, q = (uint32_t)uecrt::tc_quad();
    for (uint64_t i = 0; i < n_rays; ++i)
      for (uint32_t j = 0; j < n; ++j) {
        const uint32_t bit = j / q, pass = bit / (uint32_t)bits_per_pass, b = bit % (uint32_t)bits_per_pass;
        const uint32_t w = words[i * words_per_ray + pass * (bits_per_pass / 32) + b / 32];
        flags_out[i * n + j] = (w >> (31u - (b & 31u))) & 1u;
      }
  }
  ctx_free(c);
  return rc;
}

int uecrt_primary_hits(const uecrt_scene* scene, uint64_t n_rays, const float* o_xyz, const float* d_xyz, double t_min,
                       double t_max, int exact, double* t, double* n_xyz, int32_t* id, uint8_t* front_face) {
  int rc = require_device();
  if (rc) return rc;
  if (n_rays == 0) return UECRT_OK;
  uecrt_ctx* c = nullptr;
  if ((rc = ctx_create_on_current_device(scene, nullptr, &c))) return rc;
  rc = ctx_primary_hits(c, n_rays, o_xyz, d_xyz, t_min, t_max, exact, t, n_xyz, id, front_face);
  ctx_free(c);
  return rc;
}

int uecrt_scatter(uint64_t n, const uint32_t* mat_type, const uecrt_float4* mat_param, const float* d_in_xyz,
                  const float* p_xyz, const float* normal_xyz, const uint8_t* front_face, const float* uniforms4,
                  uint8_t* scattered, float* attenuation_rgb, float* o_out_xyz, float* d_out_xyz) {
  int rc = require_device();
  if (rc) return rc;
  if (n == 0) return UECRT_OK;
  void* bufs[12] = {};
  const size_t sz[12] = {n * 4, n * 16, n * 12, n * 12, n * 12, n, n * 16, n, n * 12, n * 12, n * 12, 0};
  const void* src[7] = {mat_type, mat_param, d_in_xyz, p_xyz, normal_xyz, front_face, uniforms4};
  auto body = [&]() -> int {
    for (int i = 0; i < 11; ++i) CUDA_TRY(cudaMalloc(&bufs[i], sz[i]));
    for (int i = 0; i < 7; ++i) CUDA_TRY(cudaMemcpy(bufs[i], src[i], sz[i], cudaMemcpyHostToDevice));
    uecrt::launch_scatter(n, (const uint32_t*)bufs[0], (const float4*)bufs[1], (const float*)bufs[2], (const float*)bufs[3],
                          (const float*)bufs[4], (const uint8_t*)bufs[5], (const float*)bufs[6], (uint8_t*)bufs[7],
                          (float*)bufs[8], (float*)bufs[9], (float*)bufs[10], 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(scattered, bufs[7], sz[7], cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(attenuation_rgb, bufs[8], sz[8], cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(o_out_xyz, bufs[9], sz[9], cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(d_out_xyz, bufs[10], sz[10], cudaMemcpyDeviceToHost));
    return UECRT_OK;
  };
  rc = body();
  for (void* b : bufs) cudaFree(b);
  return rc;
}

int uecrt_philox(uint64_t n, uint64_t seed, const uint32_t* counters4, uint32_t* out4) {
  int rc = require_device();
  if (rc) return rc;
  if (n == 0) return UECRT_OK;
  uint32_t *d_c = nullptr, *d_o = nullptr;
  auto body = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_c, n * 16)); CUDA_TRY(cudaMalloc(&d_o, n * 16));
    CUDA_TRY(cudaMemcpy(d_c, counters4, n * 16, cudaMemcpyHostToDevice));
    uecrt::launch_philox(n, (uint32_t)(seed & 0xffffffffull), (uint32_t)(seed >> 32), d_c, d_o, 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out4, d_o, n * 16, cudaMemcpyDeviceToHost));
    return UECRT_OK;
  };
  rc = body();
  cudaFree(d_c); cudaFree(d_o);
  return rc;
}

}  // extern "C"
