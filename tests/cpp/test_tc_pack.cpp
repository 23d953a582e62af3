// Host-side check of the tensor-scan operand packing (csrc/scene_pack.hpp pack_scene_tc, csrc/tc_scan.cuh
// fill_row_a / fill_row_b / split_tf32) -- no GPU.  For rays and spheres of the RTIOW final scene it
// emulates what the tensor core is given (TF32-exact hi / lo slots, exact products, exact sum) and checks
//   1. every slot value is exactly representable in TF32 (low 13 mantissa bits zero);
//   2. the canonical-layout byte offsets place row r, chunk c where the UMMA descriptor expects them;
//   3. the conservative-superset property: whenever the reference's double-precision discriminant
//      (sphere.hpp:27-32) is >= 0, the emulated accumulator G = a.b is > 0 by at least the part of
//      the bound reserved for the hardware's accumulation error (80 u S).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../uecraytracing_b200/csrc/scene_pack.hpp"

using namespace uecrt;

static bool tf32_exact(float v) { uint32_t b; std::memcpy(&b, &v, 4); return (b & 0x1fffu) == 0; }

int main() {
  std::mt19937 gen(7);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  // final-scene-like world: ground r=1000, three r=1 spheres, r=0.2 spheres on the [-11,11]^2 grid
  std::vector<uecrt_float4> geom;
  geom.push_back({0.f, -1000.f, 0.f, 1000.f});
  for (int a = -11; a < 11; ++a)
    for (int b = -11; b < 11; ++b) geom.push_back({(float)(a + 0.9 * U(gen)), 0.2f, (float)(b + 0.9 * U(gen)), 0.2f});
  geom.push_back({0.f, 1.f, 0.f, 1.f}); geom.push_back({-4.f, 1.f, 0.f, 1.f}); geom.push_back({4.f, 1.f, 0.f, 1.f});
  std::vector<uint32_t> mt(geom.size(), 0); std::vector<uecrt_float4> mp(geom.size(), uecrt_float4{.5f, .5f, .5f, 0.f});
  uecrt_scene sc{(uint32_t)geom.size(), geom.data(), mt.data(), mp.data()};
  const PackedSceneTc p = pack_scene_tc(sc, 128);
  if (p.rows % 128 || p.rows < sc.n || p.bytes.size() != (size_t)p.rows * 128) { std::printf("bad row count\n"); return 1; }
  // 1 + 2: slots TF32-exact, layout offsets
  std::vector<float> brow((size_t)p.rows * tc::kSlots);
  for (uint32_t r = 0; r < p.rows; ++r)
    for (uint32_t ch = 0; ch < 8; ++ch) {
      const uint32_t off = tc::row_chunk_offset(r, ch);
      if (off != (r / 8) * 1024 + ch * 128 + (r % 8) * 16) { std::printf("layout offset\n"); return 1; }
      std::memcpy(&brow[(size_t)r * tc::kSlots + ch * 4], &p.bytes[off], 16);
    }
  for (float v : brow) if (!tf32_exact(v)) { std::printf("slot not TF32-exact\n"); return 1; }
  for (uint32_t r = sc.n; r < p.rows; ++r) if (brow[(size_t)r * tc::kSlots + 9] > -1e29f) { std::printf("pad row can hit\n"); return 1; }
  // 2b: the swept form (scenes above the residency limit) pads to whole 512-row slices, and a slice is one contiguous
  // 64 KB block (rows are stored in 8-row groups of 1024 B), so the kernel can bulk-copy slice p from byte p * 65536
  {
    const PackedSceneTc q = pack_scene_tc(sc, 512);
    if (q.rows % 512 || q.rows < sc.n || q.bytes.size() != (size_t)q.rows * 128) { std::printf("bad swept row count\n"); return 1; }
    for (uint32_t r = 0; r < q.rows; ++r) {
      const uint32_t lo = tc::row_chunk_offset(r, 0), hi = tc::row_chunk_offset(r, 7) + 16;
      if (lo / 65536 != r / 512 || (hi - 1) / 65536 != r / 512) { std::printf("row %u leaves its slice\n", r); return 1; }
    }
    if (std::memcmp(q.bytes.data(), p.bytes.data(), p.bytes.size()) != 0) { std::printf("swept packing differs on the shared rows\n"); return 1; }
    float last[4];
    std::memcpy(last, &q.bytes[tc::row_chunk_offset(q.rows - 1, 2)], 16);        // slots 8..11: slot 9 = hi(-k + bound)
    if (last[1] > -1e29f) { std::printf("swept pad row can hit\n"); return 1; }
  }
  // 3: conservative superset on random rays (camera rays and rays leaving surfaces)
  const double u = std::ldexp(1.0, -24);
  long tested = 0, hits = 0;
  double min_margin = 1e300;
  for (int it = 0; it < 20000; ++it) {
    float ox, oy, oz;
    if (it & 1) { ox = 13.f; oy = 2.f; oz = 3.f; }
    else { ox = (float)(-40 + 80 * U(gen)); oy = (float)(U(gen) < .5 ? 0.0 : 0.4 * U(gen)); oz = (float)(-40 + 80 * U(gen)); }
    double tx = (it & 1) ? (-11 + 22 * U(gen)) - ox : 2 * U(gen) - 1, ty = (it & 1) ? (0.2 - oy) + (U(gen) - .5) : 2 * U(gen) - 1,
           tz = (it & 1) ? (-11 + 22 * U(gen)) - oz : 2 * U(gen) - 1;
    // the device's ray: FP32 normalisation, then make_scan_ray's FP32 arithmetic (closest_hit.cuh)
    const float len2 = std::fmaf((float)tz, (float)tz, std::fmaf((float)ty, (float)ty, (float)tx * (float)tx));
    const float inv = 1.0f / std::sqrt(len2);
    const float dx = (float)tx * inv, dy = (float)ty * inv, dz = (float)tz * inv;
    const float od = std::fmaf(oz, dz, std::fmaf(oy, dy, ox * dx));
    const float nqx = 2.0f * std::fmaf(od, dx, -ox), nqy = 2.0f * std::fmaf(od, dy, -oy), nqz = 2.0f * std::fmaf(od, dz, -oz);
    const float oo = std::fmaf(oz, oz, std::fmaf(oy, oy, ox * ox));
    const float P = std::fmaf(-od, od, oo) - oo * (208.0f * 5.9604644775390625e-8f);
    const float a[tc::kTerms] = {dx * dx, dy * dy, dz * dz, 2.0f * dx * dy, 2.0f * dx * dz, 2.0f * dy * dz, nqx, nqy, nqz, 1.0f};
    float arow[tc::kSlots];
    tc::fill_row_a(a, P, arow);
    for (float v : arow) if (!tf32_exact(v)) { std::printf("ray slot not TF32-exact\n"); return 1; }
    for (uint32_t n = 0; n < sc.n; n += 1 + (it % 3)) {
      // the reference's discriminant for the ORIGINAL (unnormalised) direction, in double
      const double cx = geom[n].x, cy = geom[n].y, cz = geom[n].z, r = geom[n].w;
      const double Dx = (float)tx, Dy = (float)ty, Dz = (float)tz;
      const double ocx = ox - cx, ocy = oy - cy, ocz = oz - cz;
      const double aa = Dx * Dx + Dy * Dy + Dz * Dz, hb = ocx * Dx + ocy * Dy + ocz * Dz, c = ocx * ocx + ocy * ocy + ocz * ocz - r * r;
      const double disc = hb * hb - aa * c;
      double G = 0, S = 0;
      for (int k = 0; k < tc::kSlots; ++k) { const double t = (double)arow[k] * brow[(size_t)n * tc::kSlots + k]; G += t; S += std::fabs(t); }
      ++tested;
      if (disc >= 0) {
        ++hits;
        const double margin = G - 80.0 * u * S;      // what is left after the hardware's share of the bound (truncating-adder model, scene_pack.hpp)
        if (margin < min_margin) min_margin = margin;
        if (!(margin > 0)) { std::printf("superset violated: ray %d sphere %u disc %.3e G %.3e S %.3e\n", it, n, disc, G, S); return 1; }
      }
    }
  }
  if (hits < 1000) { std::printf("too few hits (%ld of %ld)\n", hits, tested); return 1; }
  std::printf("tc pack ok: %ld pairs, %ld with disc >= 0, min margin %.3e\n", tested, hits, min_margin);
  return 0;
}
