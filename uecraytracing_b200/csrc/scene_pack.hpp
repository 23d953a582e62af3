// scene_pack.hpp -- host-side flattening of the sphere list into the scan layout.
//
// Per sphere PAIR (2i, 2i+1), 32 bytes, read in the K1/K2 scan as two broadcast LDS.128
// (the scene is cut into TILES of tile_pairs pairs; one tile when it fits in shared memory):
//     A = { -cx0, -cx1, -cy0, -cy1 }      B = { -cz0, -cz1, nk0, nk1 }
// with  nk = -(|c|^2 - r^2) + bound_sphere   (computed in double, rounded UP to float).
//
// Error bound (why the FP32 scan never drops a sphere the double reference would hit).
// With u = 2^-24, unit direction d (|d| = 1 +- 3u after FP32 normalisation), M = |o| + |c|:
//   hb   = fl(o.d - c.d)            |err| <= 10u M            -> hb^2 off by <= 21u M^2
//   cc   = fl(|o|^2 - 2o.c + k)     |err| <= 7u M^2 + 5u|k|   (k = |c|^2 - r^2, rounded once)
//   disc = fl(hb^2 - cc)            one more rounding <= 3u M^2
//   => |disc_fp32 - disc_exact| <= 31u M^2 + 5u|k| <= 62u|o|^2 + 62u|c|^2 + 5u(|c|^2 + r^2)
// The ray part (62u|o|^2, applied as 208u: kRayBound is shared with the tensor scan, which needs 204u) lowers |o|^2 in
// make_scan_ray (closest_hit.cuh);
// the sphere part (applied as 88u|c|^2 + 10u r^2) is folded into nk here.  Both only make
// disc_fp32 larger, so {disc_fp32 >= 0} is a superset of {disc_reference >= 0}; the resolve
// phase re-tests every candidate, so false positives cost time, never correctness.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/uecrt.h"
#include "tc_scan.cuh"

namespace uecrt {

constexpr int kPairUnroll = 8;   // pairs per unrolled scan trip (2 iterations x 2 groups x 2 pairs); npairs is padded to a multiple
constexpr int kPairPad = 2;      // one extra never-hit group after the last, read by the scan's prefetch

struct PackedScene {
  uint32_t n = 0;           // spheres
  uint32_t tile_pairs = 0;  // pairs the scan loops over per tile (multiple of kPairUnroll)
  uint32_t ntiles = 0;      // 1 = the whole scene is one resident tile
  std::vector<float> pairs; // ntiles * (tile_pairs + kPairPad) * 8 floats; every tile carries its own prefetch pad
  size_t tile_bytes() const { return (size_t)(tile_pairs + kPairPad) * 32; }
  size_t bytes() const { return pairs.size() * sizeof(float); }
};

inline float round_up_to_float(double v) {
  float f = (float)v;
  if ((double)f < v) f = std::nextafterf(f, INFINITY);
  return f;
}

// tile_spheres == 0: one tile holding everything.  Otherwise tiles of that many spheres (rounded up
// to a whole number of unrolled scan trips); sphere i lives in tile i / (2*tile_pairs).
inline PackedScene pack_scene(const uecrt_scene& sc, uint32_t tile_spheres = 0) {
  PackedScene p;
  p.n = sc.n;
  uint32_t np = (sc.n + 1) / 2;
  np = (np + kPairUnroll - 1) / kPairUnroll * kPairUnroll;
  if (np == 0) np = kPairUnroll;
  if (tile_spheres == 0 || (tile_spheres + 1) / 2 >= np) {
    p.tile_pairs = np;
    p.ntiles = 1;
  } else {
    uint32_t tp = (tile_spheres + 1) / 2;
    tp = (tp + kPairUnroll - 1) / kPairUnroll * kPairUnroll;
    p.tile_pairs = tp;
    p.ntiles = (np + tp - 1) / tp;
  }
  const uint32_t stride = p.tile_pairs + kPairPad;
  p.pairs.assign((size_t)p.ntiles * stride * 8, 0.0f);
  // default-fill with never-hit pairs: k = +1e30 -> f = -1e30 < P for every ray
  for (size_t q = 0; q < (size_t)p.ntiles * stride; ++q) { p.pairs[q * 8 + 6] = -1e30f; p.pairs[q * 8 + 7] = -1e30f; }
  const double u = 5.9604644775390625e-8;   // 2^-24
  for (uint32_t i = 0; i < sc.n; ++i) {
    const double cx = sc.geom[i].x, cy = sc.geom[i].y, cz = sc.geom[i].z, r = sc.geom[i].w;
    const double c2 = cx * cx + cy * cy + cz * cz;
    const double k = c2 - r * r;
    const double bound = 88.0 * u * c2 + 10.0 * u * r * r;
    const uint32_t pair = i / 2, half = i & 1;
    const uint32_t tile = pair / p.tile_pairs, within = pair % p.tile_pairs;
    float* q = &p.pairs[((size_t)tile * stride + within) * 8];
    q[0 + half] = -sc.geom[i].x; q[2 + half] = -sc.geom[i].y; q[4 + half] = -sc.geom[i].z;
    q[6 + half] = round_up_to_float(-k + bound);
  }
  return p;
}

// ---------------------------------------------------------------------------------------
// Tensor-core scan operand (tc_scan.cuh): one 32-slot row per sphere, rows padded to a multiple
// of `chunk` with never-hit rows, stored in the UMMA K-major no-swizzle canonical layout.
//   b = [cx^2, cy^2, cz^2, cx cy, cx cz, cy cz, -cx, -cy, -cz, nk],  nk = -(|c|^2 - r^2) + bound
// (the ray side carries -q in slots 6..8, so the product is +q.c).
//
// Error bound of the tensor-core evaluation of G = f - P against the exact value on the same float inputs.
// u = 2^-24; S = sum over the product's terms of |a_i b_i| <= 2|o|^2 + 3|c|^2 + r^2
//   (the quadratic part sums to (|cx dx| + |cy dy| + |cz dz|)^2 <= |c|^2, |q.c| <= 2|o||c| <= |o|^2 + |c|^2,
//    |k| <= |c|^2 + r^2, |P| <= |o|^2).
//   (1) three-way TF32 split (tc_scan.cuh split_tf32, lo rounded to nearest): v = hi + lo + e, |e| <= 4u|v| on both
//       operands, and the product drops a_lo b_lo <= 2^-22 |a b| = 4u|a b|                     <= 12 u S
//   (2) accumulation in the tensor core.  The 32 products of TF32 values are exact in FP32; the hardware adds them to the
//       FP32 accumulator in four K = 8 steps.  What it does inside a step is not documented, so the bound assumes the worst
//       of the known designs, a TRUNCATING aligned adder: the eight products and the accumulator are aligned to the largest
//       of them, bits below its last place (2u relative) are dropped from every addend, and the sum is truncated to FP32:
//       <= (9 + 1) x 2u x max(|addend|, |sum|) <= 20 u S per step                                <= 80 u S
//       Measured (tools/microbench/tc_probe.cu T5, profiles/tc_probe_r2_accumulate.log: 2.6e7 pairs of TF32-exact
//       operands in 25 adversarial family pairs -- same sign, one dominant slot, full cancellation, geometric spread,
//       random): 15.2 u S at most, results below the exact sum twice as often as above (partly truncating); the scene-
//       shaped probe T2 of round 1 saw 11.9 u S including the split.  The model bound covers the worst measurement 5x.
//   (3) rounding of the lifted ray vector      <= u|c|^2 + 4u(|o|^2 + |c|^2)
//       rounding of the lifted sphere vector   <= 1.5 u|c|^2
//       rounding of P                          <= 4 u|o|^2
//       direction not exactly unit (|d|^2 = 1 +- 6u) <= 12u(|o|^2 + |c|^2) + 6u r^2
//   => |G_tc - G_exact| <= 92 u S + 20u|o|^2 + 19u|c|^2 + 7u r^2 <= 204u|o|^2 + 295u|c|^2 + 99u r^2
// applied as 208u|o|^2 on the ray (kRayBound, closest_hit.cuh) and 304u|c|^2 + 104u r^2 here.  Round 1 budgeted (1) + (2)
// as 24 u S from the scene-shaped measurement alone; the GPU test tests/test_gpu_tensor_bound.py reads the scan's verdict
// back and checks the superset property on millions of (ray, sphere) pairs that sit on the decision boundary.
// Far from the origin the margins (proportional to |c|^2, |o|^2) exceed r^2 and the scan flags everything: still a superset,
// the FP32 re-test and the resolve then do the work (a scene 3000 units out renders correctly, at FP32-scan speed).
struct PackedSceneTc {
  uint32_t n = 0, rows = 0;              // spheres, padded rows (multiple of chunk)
  std::vector<unsigned char> bytes;      // rows * 128
};

inline PackedSceneTc pack_scene_tc(const uecrt_scene& sc, uint32_t chunk) {
  PackedSceneTc p;
  p.n = sc.n;
  p.rows = (sc.n + chunk - 1) / chunk * chunk;
  if (p.rows == 0) p.rows = chunk;
  p.bytes.assign((size_t)p.rows * 128, 0);
  const double u = 5.9604644775390625e-8;
  for (uint32_t i = 0; i < p.rows; ++i) {
    float b[tc::kTerms] = {0, 0, 0, 0, 0, 0, 0, 0, 0, -1e30f};   // never-hit pad row
    if (i < sc.n) {
      const double cx = sc.geom[i].x, cy = sc.geom[i].y, cz = sc.geom[i].z, r = sc.geom[i].w;
      const double c2 = cx * cx + cy * cy + cz * cz;
      const double bound = 304.0 * u * c2 + 104.0 * u * r * r;
      b[0] = (float)(cx * cx); b[1] = (float)(cy * cy); b[2] = (float)(cz * cz);
      b[3] = (float)(cx * cy); b[4] = (float)(cx * cz); b[5] = (float)(cy * cz);
      b[6] = -sc.geom[i].x; b[7] = -sc.geom[i].y; b[8] = -sc.geom[i].z;
      b[9] = round_up_to_float(-(c2 - r * r) + bound);
    }
    float row[tc::kSlots];
    tc::fill_row_b(b, row);
    for (uint32_t ch = 0; ch < 8; ++ch) std::memcpy(&p.bytes[tc::row_chunk_offset(i, ch)], &row[4 * ch], 16);
  }
  return p;
}

// Quad table of the tensor scan's shade side (tc_collect.cuh quad_candidates): for spheres 4q .. 4q+3 four float4
// {-cx[4]}, {-cy[4]}, {-cz[4]}, {nk[4]} with the FP32 scan's nk (same value, same bound as pack_scene above), so a flagged
// quad is re-tested with exactly the FP32 scan's arithmetic.  Rows beyond n are never-hit (nk = -1e30).
inline std::vector<float> pack_scene_quads(const uecrt_scene& sc, uint32_t rows) {
  std::vector<float> q((size_t)rows * 4, 0.0f);
  const double u = 5.9604644775390625e-8;   // 2^-24
  for (uint32_t i = 0; i < rows; ++i) {
    float* base = &q[(size_t)(i / 4) * 16 + (i & 3)];
    base[12] = -1e30f;
    if (i >= sc.n) continue;
    const double cx = sc.geom[i].x, cy = sc.geom[i].y, cz = sc.geom[i].z, r = sc.geom[i].w;
    const double c2 = cx * cx + cy * cy + cz * cz;
    const double bound = 88.0 * u * c2 + 10.0 * u * r * r;
    base[0] = -sc.geom[i].x; base[4] = -sc.geom[i].y; base[8] = -sc.geom[i].z;
    base[12] = round_up_to_float(-(c2 - r * r) + bound);
  }
  return q;
}

}  // namespace uecrt
