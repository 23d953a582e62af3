// kernels.cuh -- kernel argument blocks shared by kernels.cu (device) and uecrt_api.cu (host).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uecrt {

// uniform grid over the spheres (grid_build.hpp): per level, cell (x, y, z) lists items[cell_start[c] .. cell_start[c+1]),
// c = (z * dim[1] + y) * dim[0] + x.  Two levels (fine: the typical spheres; coarse: the ones too large for it, empty when
// there are only a few); `always` spheres are tested for every ray.
struct DeviceGridLevel {
  float org[3], cell[3], inv_cell[3];
  int dim[3];
  const uint32_t* cell_start;     // null: level not present
  const uint32_t* items;
};
struct DeviceGrid {
  DeviceGridLevel lvl[2];         // [0] fine, [1] coarse
  const uint32_t* always;
  int n_always;
};

struct DeviceScene {
  const ulonglong2* pairs;   // global copy of the packed pairs (2 x 16 B per pair), staged to smem by TMA
  int npairs;                // pairs per TILE (multiple of 8); every tile is stored with npairs + 2 (prefetch pad)
  int ntiles;                // 1: the whole scene is resident in shared memory; >1: swept through a 2-buffer ring
  int n;                     // spheres
  const float4* geom;        // {cx,cy,cz,r}
  const uint32_t* mat_type;
  const float4* mat_param;
  const unsigned char* tc_rows;   // tensor-core scan operand (scene_pack.hpp pack_scene_tc), or null
  int tc_nrows;                   // rows (spheres padded to a multiple of tc::kNC, or of the pass size when swept)
  int tc_npasses;                 // 1: the operand is resident in shared memory; > 1: swept in that many equal slices
  const float4* tc_quads;         // per 4 spheres {-cx[4]}, {-cy[4]}, {-cz[4]}, {nk[4]} (scene_pack.hpp pack_scene_quads), tc_nrows / 4 quads
  DeviceGrid grid;
};

struct DeviceCamera {
  float origin[3], ll[3], hor[3], ver[3], u[3], v[3];
  float lens;
  int origin_is_zero;        // the reference never subtracts its zero origin (camera.hpp:29-32)
};

// counters[0] work-queue head, [1] rays traced, [2] paths finished
struct RenderArgs {
  DeviceScene sc;
  DeviceCamera cam;
  uint32_t width, height;
  uint32_t sample_begin, sample_count;          // global sample ids [begin, begin+count) of every pixel
  uint32_t max_depth;
  uint32_t key0, key1;                          // Philox key (seed)
  uint32_t tiles_x;                             // 8x4-pixel tiles per row (ceil(width / 8))
  float inv_tiles_x;
  unsigned long long n_slots;                   // tiles * 32 (pixel slots incl. the ragged edge)
  unsigned long long total_items;               // work items of this launch (one per path slot): [item_begin, item_begin + total_items)
  unsigned long long item_begin;                //   of the n_slots * sample_count items of the sample range, sample-major
  float4* accum;
  unsigned long long* counters;
  int count_stats;
};

constexpr int kRenderThreads = 256;
constexpr int kHitsThreads = 128;

int max_rays_per_thread();   // 2 only when built with -DUECRT_ENABLE_RAYS2
size_t render_smem_bytes(int npairs, int ntiles, int threads, int rays);

void launch_render_paths(const RenderArgs& a, bool exact, int rays, int blocks, int threads, size_t smem, cudaStream_t st);
void launch_primary_hits(const DeviceScene& sc, unsigned long long n_rays, const float* o, const float* d,
                         double tmin, double tmax, bool exact, double* t, double* nrm, int32_t* id,
                         uint8_t* ff, cudaStream_t st);
void launch_resolve(const float4* accum, uint32_t n_pixels, uint32_t spp, uint8_t* rgb, cudaStream_t st);
void launch_scatter(unsigned long long n, const uint32_t* mat_type, const float4* mat_param, const float* d_in,
                    const float* p, const float* nrm, const uint8_t* ff, const float* uniforms4,
                    uint8_t* scattered, float* att, float* o_out, float* d_out, cudaStream_t st);
void launch_camera_rays(const RenderArgs& a, unsigned long long n, const uint32_t* pixel, const uint32_t* sample, float* o, float* d, cudaStream_t st);
void launch_philox(unsigned long long n, uint32_t k0, uint32_t k1, const uint32_t* ctr4, uint32_t* out4, cudaStream_t st);
cudaError_t configure_kernels(int* regs_render, int* max_blocks_per_sm, size_t smem_render, bool tiled, int rays);

// ---- uniform-grid variants (grid_walk.cuh): counters[3] receives the sphere tests actually executed
void launch_render_paths_grid(const RenderArgs& a, bool exact, int blocks, cudaStream_t st);
void launch_primary_hits_grid(const DeviceScene& sc, unsigned long long n_rays, const float* o, const float* d,
                              double tmin, double tmax, bool exact, double* t, double* nrm, int32_t* id,
                              uint8_t* ff, cudaStream_t st);
cudaError_t configure_kernels_grid(int* regs_render, int* max_blocks_per_sm);

// ---- tensor-core scan variants (tc_scan.cuh): the sphere operand is resident in shared memory up to
// tc_resident_max_spheres() spheres and swept through it in slices of tc_sweep_rows(n) rows above that
int tc_resident_max_spheres();
int tc_sweep_rows(int n_spheres);
int tc_threads();
int tc_chunk();
size_t tc_smem_bytes(int tc_nrows, int tc_npasses);
cudaError_t configure_kernels_tc(int* regs_render, size_t smem);
void launch_render_paths_tc(const RenderArgs& a, bool exact, int blocks, size_t smem, cudaStream_t st);
void launch_primary_hits_tc(const DeviceScene& sc, unsigned long long n_rays, const float* o, const float* d,
                            double tmin, double tmax, bool exact, double* t, double* nrm, int32_t* id,
                            uint8_t* ff, int sm_count, cudaStream_t st, uint32_t* dbg_words = nullptr);
int tc_quad();               // spheres per candidate-bitmap bit of the tensor scan (1 or 4)
// K1 launch shape for a tensor scan of this scene; pipeline: 1 = self-scan (tc_self.cuh), 0 = scan warps (tc_scan.cuh)
void tc_render_config(const DeviceScene& sc, int* threads, size_t* smem, int* pipeline);
void tc_set_pipeline(int pipeline);

}  // namespace uecrt
