// Host-side check of the uniform grid (csrc/grid_build.hpp) -- no GPU.  The invariant the device walk relies
// on: every point of every sphere's surface lies in a cell that lists the sphere (or the sphere is in the
// always-list), cell lists are ascending, and the CSR directory is consistent.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../../uecraytracing_b200/csrc/grid_build.hpp"

using namespace uecrt;

static int check(const std::vector<uecrt_float4>& geom, const char* what) {
  uecrt_scene sc{(uint32_t)geom.size(), geom.data(), nullptr, nullptr};
  const HostGrid g = build_grid(sc);
  const size_t ncell = (size_t)g.dim[0] * g.dim[1] * g.dim[2];
  if (!geom.empty() && !g.items.empty() && g.cell_start.size() != ncell + 1) { std::printf("%s: directory size\n", what); return 1; }
  for (size_t c = 0; c + 1 < g.cell_start.size(); ++c) {
    if (g.cell_start[c] > g.cell_start[c + 1]) { std::printf("%s: directory not monotone\n", what); return 1; }
    for (uint32_t k = g.cell_start[c]; k + 1 < g.cell_start[c + 1]; ++k)
      if (g.items[k] >= g.items[k + 1]) { std::printf("%s: cell list not ascending\n", what); return 1; }
  }
  if (!g.cell_start.empty() && g.cell_start.back() != g.items.size()) { std::printf("%s: directory end\n", what); return 1; }
  std::vector<char> always(geom.size(), 0);
  for (uint32_t i : g.always) always[i] = 1;
  std::mt19937 gen(3);
  std::normal_distribution<double> N(0.0, 1.0);
  long pts = 0;
  for (uint32_t i = 0; i < geom.size(); ++i) {
    if (always[i]) continue;
    for (int s = 0; s < 24; ++s) {
      double v[3] = {N(gen), N(gen), N(gen)};
      const double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), r = std::fabs(geom[i].w);
      const double p[3] = {geom[i].x + r * v[0] / l, geom[i].y + r * v[1] / l, geom[i].z + r * v[2] / l};
      int idx[3];
      for (int a = 0; a < 3; ++a) {
        idx[a] = (int)std::floor(((float)p[a] - g.org[a]) * g.inv_cell[a]);       // the device's FP32 cell coordinate
        if (idx[a] < 0 || idx[a] >= g.dim[a]) { std::printf("%s: surface point of sphere %u outside the grid\n", what, i); return 1; }
      }
      const size_t c = ((size_t)idx[2] * g.dim[1] + idx[1]) * g.dim[0] + idx[0];
      bool found = false;
      for (uint32_t k = g.cell_start[c]; k < g.cell_start[c + 1]; ++k) found = found || g.items[k] == i;
      if (!found) { std::printf("%s: sphere %u missing from the cell of one of its surface points\n", what, i); return 1; }
      ++pts;
    }
  }
  std::printf("%s: %zu spheres, grid %d x %d x %d, %zu list entries (%.2f per sphere), %zu always, %ld surface points ok\n", what,
              geom.size(), g.dim[0], g.dim[1], g.dim[2], g.items.size(), geom.empty() ? 0.0 : (double)g.items.size() / geom.size(),
              g.always.size(), pts);
  return 0;
}

// the two-level form: every sphere is in the always-list or every point of its surface lies in a cell of ITS level that lists it
static int in_level(const HostGrid& g, uint32_t i, const double p[3]) {
  if (g.items.empty()) return 0;
  int idx[3];
  for (int a = 0; a < 3; ++a) {
    idx[a] = (int)std::floor(((float)p[a] - g.org[a]) * g.inv_cell[a]);
    if (idx[a] < 0 || idx[a] >= g.dim[a]) return 0;
  }
  const size_t c = ((size_t)idx[2] * g.dim[1] + idx[1]) * g.dim[0] + idx[0];
  for (uint32_t k = g.cell_start[c]; k < g.cell_start[c + 1]; ++k) if (g.items[k] == i) return 1;
  return 0;
}
static int check_levels(const std::vector<uecrt_float4>& geom, const char* what, bool want_coarse) {
  uecrt_scene sc{(uint32_t)geom.size(), geom.data(), nullptr, nullptr};
  const HostGrids G = build_grids(sc);
  std::vector<char> where(geom.size(), 0);      // 1 always, 2 fine, 4 coarse (by membership anywhere in that level)
  for (uint32_t i : G.always) where[i] |= 1;
  for (uint32_t i : G.fine.items) where[i] |= 2;
  for (uint32_t i : G.coarse.items) where[i] |= 4;
  std::mt19937 gen(5);
  std::normal_distribution<double> N(0.0, 1.0);
  for (uint32_t i = 0; i < geom.size(); ++i) {
    if (where[i] != 1 && where[i] != 2 && where[i] != 4) { std::printf("%s: sphere %u is in %d places\n", what, i, (int)where[i]); return 1; }
    if (where[i] == 1) continue;
    for (int s = 0; s < 24; ++s) {
      double v[3] = {N(gen), N(gen), N(gen)};
      const double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), r = std::fabs(geom[i].w);
      const double p[3] = {geom[i].x + r * v[0] / l, geom[i].y + r * v[1] / l, geom[i].z + r * v[2] / l};
      if (!in_level(where[i] == 2 ? G.fine : G.coarse, i, p)) { std::printf("%s: sphere %u missing from its level's cell\n", what, i); return 1; }
    }
  }
  if (want_coarse != !G.coarse.items.empty()) { std::printf("%s: coarse level %s\n", what, want_coarse ? "missing" : "unexpected"); return 1; }
  std::printf("%s: fine %d x %d x %d with %zu entries, coarse %d x %d x %d with %zu entries, %zu always (%.2f entries per sphere)\n", what, G.fine.dim[0], G.fine.dim[1],
              G.fine.dim[2], G.fine.items.size(), G.coarse.dim[0], G.coarse.dim[1], G.coarse.dim[2], G.coarse.items.size(), G.always.size(),
              geom.empty() ? 0.0 : (double)(G.fine.items.size() + G.coarse.items.size()) / geom.size());
  return 0;
}

int main() {
  std::mt19937 gen(11);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  std::vector<uecrt_float4> dense, fin, far, one, none;
  dense.push_back({0.f, -1000.f, 0.f, 1000.f});
  for (int i = 0; i < 10000; ++i) { const float r = 0.1f + 0.3f * U(gen); dense.push_back({-50.f + 100.f * U(gen), r, -50.f + 100.f * U(gen), r}); }
  fin.push_back({0.f, -1000.f, 0.f, 1000.f});
  for (int a = -11; a < 11; ++a) for (int b = -11; b < 11; ++b) fin.push_back({a + 0.9f * U(gen), 0.2f, b + 0.9f * U(gen), 0.2f});
  fin.push_back({0.f, 1.f, 0.f, 1.f}); fin.push_back({-4.f, 1.f, 0.f, 1.f}); fin.push_back({4.f, 1.f, 0.f, -1.f});   // a negative radius too
  for (const auto& s : fin) far.push_back({s.x + 3000.f, s.y + 50.f, s.z - 2000.f, s.w});
  one.push_back({0.f, 0.f, -1.f, 0.5f});
  int rc = check(dense, "dense") | check(fin, "final-like") | check(far, "far from origin") | check(one, "one sphere") | check(none, "empty");
  // 3-D cloud with very different radii
  std::vector<uecrt_float4> cloud;
  for (int i = 0; i < 3000; ++i) cloud.push_back({-6.f + 12.f * U(gen), -6.f + 12.f * U(gen), -6.f + 12.f * U(gen), 0.02f + (i % 97 == 0 ? 3.f : 0.5f) * U(gen)});
  rc |= check(cloud, "cloud");
  // two levels: 8000 pebbles, 200 boulders 20 - 40 x their size, the ground
  std::vector<uecrt_float4> mixed;
  mixed.push_back({0.f, -1000.f, 0.f, 1000.f});
  for (int i = 0; i < 8000; ++i) { const float r = 0.1f + 0.2f * U(gen); mixed.push_back({-60.f + 120.f * U(gen), r, -60.f + 120.f * U(gen), r}); }
  for (int i = 0; i < 200; ++i) { const float r = 3.f + 3.f * U(gen); mixed.push_back({-60.f + 120.f * U(gen), r, -60.f + 120.f * U(gen), r}); }
  rc |= check_levels(mixed, "mixed radii", true) | check_levels(dense, "dense (levels)", false) | check_levels(fin, "final-like (levels)", false) |
        check_levels(none, "empty (levels)", false) | check_levels(one, "one sphere (levels)", false);
  if (!rc) std::printf("grid build ok\n");
  return rc;
}
