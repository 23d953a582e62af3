// grid_build.hpp -- host-side uniform grid over the sphere list: the spatial index behind the same
// closest-hit contract.  hittable_list::hit_impl (hittable_list.hpp:32-58) tests
// every object; the grid only changes WHICH spheres a ray tests, never the test (sphere.hpp:25-48) nor
// the result: a sphere is listed in every cell its (slightly inflated) bounding box overlaps, so every
// point where a ray can hit it lies in a cell that lists it, and spheres too large for the grid are
// tested for every ray.  The later-index-wins tie rule is kept by an order-independent comparison in
// the resolve (closest_hit.cuh test_fast / test_exact with ORDERED = false).
//
// This changes the ALGORITHMIC work (it is no longer 18 FLOP x N x rays), so it is opt-in
// (UECRT_FLAG_SCAN_GRID) and reported separately from the brute-force roofline figure.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/uecrt.h"

namespace uecrt {

struct HostGrid {
  float org[3] = {0, 0, 0};        // lower corner of the grid
  float cell[3] = {1, 1, 1};       // cell edge lengths
  float inv_cell[3] = {1, 1, 1};
  int dim[3] = {1, 1, 1};
  std::vector<uint32_t> cell_start;   // dim[0]*dim[1]*dim[2] + 1 offsets into items (x fastest, then y, then z)
  std::vector<uint32_t> items;        // sphere indices, ascending inside each cell
  std::vector<uint32_t> always;       // spheres tested for every ray (too large for the grid), ascending
};

// One grid level over the spheres `subset` (ascending indices).  Spheres of the subset whose radius exceeds 8 x the
// subset's median radius would be listed in a huge number of cells: they are returned in g.always instead.  With
// grow_to_fit (the last level) that limit is raised until at most 64 spheres are left over.
inline HostGrid build_grid_level(const uecrt_scene& sc, const std::vector<uint32_t>& subset, bool grow_to_fit) {
  HostGrid g;
  const uint32_t n = (uint32_t)subset.size();
  g.cell_start.assign(2, 0);
  if (n == 0) return g;
  std::vector<float> radii(sc.n, 0.f);
  std::vector<float> sorted;
  sorted.reserve(n);
  for (uint32_t i : subset) { radii[i] = std::fabs(sc.geom[i].w); sorted.push_back(radii[i]); }
  std::nth_element(sorted.begin(), sorted.begin() + n / 2, sorted.end());
  const float r_med = std::max(sorted[n / 2], 1e-6f);
  float big = 8.0f * r_med;
  std::vector<uint32_t> small;
  for (int attempt = 0; attempt < 8; ++attempt) {
    small.clear(); g.always.clear();
    for (uint32_t i : subset) (radii[i] <= big ? small : g.always).push_back(i);
    if (!grow_to_fit || g.always.size() <= 64) break;
    big *= 4.0f;
  }
  if (small.empty()) return g;           // everything is in the always-list: one empty cell
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, maxabs = 0;
  for (uint32_t i : small) {
    const double c[3] = {sc.geom[i].x, sc.geom[i].y, sc.geom[i].z}, r = radii[i];
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], c[a] - r); hi[a] = std::max(hi[a], c[a] + r); maxabs = std::max(maxabs, std::fabs(c[a]) + r); }
  }
  double ext[3], maxext = 0;
  for (int a = 0; a < 3; ++a) { ext[a] = hi[a] - lo[a]; maxext = std::max(maxext, ext[a]); }
  // inflation: covers the FP32 rounding of hit points and of the traversal's cell coordinates
  const double eps = 1e-4 * std::max(maxext, 1e-3) + 4e-6 * maxabs;
  for (int a = 0; a < 3; ++a) { lo[a] -= 2 * eps; hi[a] += 2 * eps; ext[a] = hi[a] - lo[a]; }
  // about 4 cells per sphere, cubic cells, never smaller than the typical diameter / 2
  double s = std::cbrt(ext[0] * ext[1] * ext[2] / (4.0 * small.size()));
  s = std::max(s, (double)r_med);
  uint64_t total = 1;
  for (int a = 0; a < 3; ++a) {
    g.dim[a] = (int)std::min(512.0, std::max(1.0, std::ceil(ext[a] / s)));
    total *= (uint64_t)g.dim[a];
  }
  while (total > (1u << 22)) {           // cap the directory at 4 M cells
    int a = (g.dim[0] >= g.dim[1] && g.dim[0] >= g.dim[2]) ? 0 : (g.dim[1] >= g.dim[2] ? 1 : 2);
    total /= (uint64_t)g.dim[a]; g.dim[a] = (g.dim[a] + 1) / 2; total *= (uint64_t)g.dim[a];
  }
  for (int a = 0; a < 3; ++a) {
    g.org[a] = (float)lo[a];
    g.cell[a] = (float)(ext[a] / g.dim[a]);
    g.inv_cell[a] = 1.0f / g.cell[a];
  }
  auto range = [&](uint32_t i, int a, int& c0, int& c1) {
    const double c = a == 0 ? sc.geom[i].x : a == 1 ? sc.geom[i].y : sc.geom[i].z, r = radii[i] + eps;
    c0 = (int)std::floor((c - r - g.org[a]) * g.inv_cell[a]);
    c1 = (int)std::floor((c + r - g.org[a]) * g.inv_cell[a]);
    c0 = std::max(0, std::min(g.dim[a] - 1, c0 - 0)); c1 = std::max(0, std::min(g.dim[a] - 1, c1));
  };
  const size_t ncell = (size_t)g.dim[0] * g.dim[1] * g.dim[2];
  std::vector<uint32_t> count(ncell + 1, 0);
  for (int pass = 0; pass < 2; ++pass) {
    for (uint32_t i : small) {           // ascending index, so every cell's list is ascending
      int x0, x1, y0, y1, z0, z1;
      range(i, 0, x0, x1); range(i, 1, y0, y1); range(i, 2, z0, z1);
      for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y)
          for (int x = x0; x <= x1; ++x) {
            const size_t c = ((size_t)z * g.dim[1] + y) * g.dim[0] + x;
            if (pass == 0) ++count[c + 1];
            else g.items[count[c]++] = i;
          }
    }
    if (pass == 0) {
      for (size_t c = 0; c < ncell; ++c) count[c + 1] += count[c];
      g.cell_start.assign(count.begin(), count.end());
      g.items.assign(count[ncell], 0);
      count.assign(g.cell_start.begin(), g.cell_start.end());
    }
  }
  return g;
}

// TWO levels: a fine grid over the typical spheres, and -- when more than kCoarseMin spheres are too large for it -- a
// coarse grid over those, sized by THEIR median radius; what is too large even for that one (the r = 1000 ground) is
// tested for every ray.  One cell size cannot serve a scene whose radii span orders of magnitude: a medium sphere in
// the fine grid is listed in thousands of cells, in the always-list it costs every ray a test.
constexpr size_t kCoarseMin = 8;
struct HostGrids {
  HostGrid fine, coarse;             // coarse.items empty: one level only
  std::vector<uint32_t> always;      // ascending
};
inline HostGrids build_grids(const uecrt_scene& sc) {
  HostGrids G;
  std::vector<uint32_t> all(sc.n);
  for (uint32_t i = 0; i < sc.n; ++i) all[i] = i;
  HostGrid one = build_grid_level(sc, all, /*grow_to_fit=*/false);
  if (one.always.size() <= kCoarseMin) {
    G.fine = build_grid_level(sc, all, true);        // a handful of large spheres: the single level with its always-list
    G.always = G.fine.always;
    G.coarse.cell_start.assign(2, 0);
  } else {
    G.coarse = build_grid_level(sc, one.always, true);
    G.always = G.coarse.always;
    G.fine = std::move(one);
  }
  G.fine.always.clear(); G.coarse.always.clear();
  return G;
}
// the single-level form (round 1), kept for its host test
inline HostGrid build_grid(const uecrt_scene& sc) {
  std::vector<uint32_t> all(sc.n);
  for (uint32_t i = 0; i < sc.n; ++i) all[i] = i;
  return build_grid_level(sc, all, true);
}

}  // namespace uecrt
