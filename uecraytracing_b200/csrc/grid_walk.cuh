// grid_walk.cuh -- device traversal of the uniform grid (grid_build.hpp): 3-D DDA (Amanatides & Woo) from
// the point where the ray enters the grid box, testing the spheres listed in every cell it passes with the
// SAME per-sphere test as the brute-force scans (closest_hit.cuh test_fast / test_exact, unordered tie rule),
// and stopping once the closest hit so far lies before the exit of the current cell.  Spheres too large
// for any level (the `always` list: e.g. the r = 1000 ground) are tested first for every ray; with two levels the coarse
// one is walked before the fine one, and the order-independent tie rule makes the order immaterial to the result.
//
// Why no hit is lost: a sphere is listed in every cell its bounding box, inflated by eps = 1e-4 x scene extent,
// overlaps; the ray's true path through the grid and the DDA's FP32 cell sequence differ only within
// ~1e-6 x coordinates of a cell boundary, where both neighbouring cells list the sphere.  The stop test keeps
// a relative margin of 1e-4 on the cell's exit distance, so candidates with (nearly) equal t in the next cell
// -- where the later-index-wins tie rule could still change the winner -- are tested too.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "kernels.cuh"

namespace uecrt {

// one level: test(j) runs the per-sphere test; best_t() returns the current closest distance along the UNIT direction
template <class Test, class BestT>
__device__ __forceinline__ void grid_walk_level(const DeviceGridLevel& G, float ox, float oy, float oz, float dx, float dy, float dz,
                                                Test&& test, BestT&& best_t, unsigned& n_tests) {
  if (!G.cell_start) return;
  const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
  // slab test against the grid box
  float t0 = 0.0f, t1 = best_t();
  float inv[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lo = G.org[a], hi = fmaf((float)G.dim[a], G.cell[a], G.org[a]);
    if (fabsf(d[a]) > 1e-20f) {
      inv[a] = 1.0f / d[a];
      const float ta = (lo - o[a]) * inv[a], tb = (hi - o[a]) * inv[a];
      t0 = fmaxf(t0, fminf(ta, tb));
      t1 = fminf(t1, fmaxf(ta, tb));
    } else {
      inv[a] = CUDART_INF_F;
      if (o[a] < lo || o[a] > hi) return;
    }
  }
  if (!(t0 <= t1)) return;
  // per axis: cell coordinate, linear-index stride (signed), steps left before the walk leaves the grid, the
  // ray parameter of the next cell boundary and its increment
  int left[3], stride[3];
  float tmax[3], tdelta[3];
  uint32_t c = 0;
  const int strides[3] = {1, G.dim[0], G.dim[0] * G.dim[1]};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float p = fmaf(d[a], t0, o[a]);
    int i = (int)floorf((p - G.org[a]) * G.inv_cell[a]);
    i = max(0, min(G.dim[a] - 1, i));
    c += (uint32_t)(i * strides[a]);
    if (fabsf(d[a]) > 1e-20f) {
      const bool fwd = d[a] > 0.0f;
      stride[a] = fwd ? strides[a] : -strides[a];
      left[a] = fwd ? G.dim[a] - 1 - i : i;
      const float boundary = fmaf((float)(i + (fwd ? 1 : 0)), G.cell[a], G.org[a]);
      tmax[a] = (boundary - o[a]) * inv[a];
      tdelta[a] = G.cell[a] * fabsf(inv[a]);
    } else {
      stride[a] = 0; left[a] = 0x7fffffff; tmax[a] = CUDART_INF_F; tdelta[a] = CUDART_INF_F;
    }
  }
  for (;;) {
    const uint32_t b = __ldg(&G.cell_start[c]), e = __ldg(&G.cell_start[c + 1]);
    for (uint32_t k = b; k < e; ++k) { test((int)__ldg(&G.items[k])); ++n_tests; }
    const float t_exit = fminf(tmax[0], fminf(tmax[1], tmax[2]));
    if (best_t() < t_exit - 1e-4f * (1.0f + fabsf(t_exit))) return;      // the closest hit lies inside the cells visited so far
    // step along the axis whose boundary comes first (x before y before z on ties), without branching: the
    // three-way branch ran at 5 - 9 active lanes (ncu source view)
    const bool sx = tmax[0] <= tmax[1] && tmax[0] <= tmax[2];
    const bool sy = !sx && tmax[1] <= tmax[2];
    const bool sz = !sx && !sy;
    left[0] -= sx ? 1 : 0; left[1] -= sy ? 1 : 0; left[2] -= sz ? 1 : 0;
    if ((left[0] | left[1] | left[2]) < 0) return;                        // left the grid
    c += (uint32_t)(sx ? stride[0] : sy ? stride[1] : stride[2]);
    tmax[0] += sx ? tdelta[0] : 0.0f; tmax[1] += sy ? tdelta[1] : 0.0f; tmax[2] += sz ? tdelta[2] : 0.0f;
  }
}

// the always-list, then the coarse level (few, large spheres: an early bound on the closest hit), then the fine level
template <class Test, class BestT>
__device__ __forceinline__ void grid_walk(const DeviceGrid& G, float ox, float oy, float oz, float dx, float dy, float dz,
                                          Test&& test, BestT&& best_t, unsigned& n_tests) {
  for (int e = 0; e < G.n_always; ++e) { test((int)__ldg(&G.always[e])); ++n_tests; }
  grid_walk_level(G.lvl[1], ox, oy, oz, dx, dy, dz, test, best_t, n_tests);
  grid_walk_level(G.lvl[0], ox, oy, oz, dx, dy, dz, test, best_t, n_tests);
}

}  // namespace uecrt
