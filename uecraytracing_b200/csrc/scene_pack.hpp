// scene_pack.hpp -- host-side flattening of the sphere list into the scan layout.
//
// Per sphere PAIR (2i, 2i+1), 32 bytes, read in the K1/K2 scan as two broadcast LDS.128:
//     A = { -cx0, -cx1, -cy0, -cy1 }      B = { -cz0, -cz1, nk0, nk1 }
// with  nk = -(|c|^2 - r^2) + bound_sphere   (computed in double, rounded UP to float).
//
// Error bound (why the FP32 scan never drops a sphere the double reference would hit).
// With u = 2^-24, unit direction d (|d| = 1 +- 3u after FP32 normalisation), M = |o| + |c|:
//   hb   = fl(o.d - c.d)            |err| <= 10u M            -> hb^2 off by <= 21u M^2
//   cc   = fl(|o|^2 - 2o.c + k)     |err| <= 7u M^2 + 5u|k|   (k = |c|^2 - r^2, rounded once)
//   disc = fl(hb^2 - cc)            one more rounding <= 3u M^2
//   => |disc_fp32 - disc_exact| <= 31u M^2 + 5u|k| <= 62u|o|^2 + 62u|c|^2 + 5u(|c|^2 + r^2)
// The ray part (62u|o|^2, applied as 80u) lowers |o|^2 in make_scan_ray (closest_hit.cuh);
// the sphere part (applied as 88u|c|^2 + 10u r^2) is folded into nk here.  Both only make
// disc_fp32 larger, so {disc_fp32 >= 0} is a superset of {disc_reference >= 0}; the resolve
// phase re-tests every candidate, so false positives cost time, never correctness.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/uecrt.h"

namespace uecrt {

constexpr int kPairUnroll = 8;   // pairs per unrolled scan trip (2 iterations x 2 groups x 2 pairs); npairs is padded to a multiple
constexpr int kPairPad = 2;      // one extra never-hit group after the last, read by the scan's prefetch

struct PackedScene {
  uint32_t n = 0;        // spheres
  uint32_t npairs = 0;   // padded pair count (multiple of kPairUnroll) the scan loops over
  std::vector<float> pairs;   // (npairs + kPairPad) * 8 floats
  size_t bytes() const { return pairs.size() * sizeof(float); }
};

inline float round_up_to_float(double v) {
  float f = (float)v;
  if ((double)f < v) f = std::nextafterf(f, INFINITY);
  return f;
}

inline PackedScene pack_scene(const uecrt_scene& sc) {
  PackedScene p;
  p.n = sc.n;
  uint32_t np = (sc.n + 1) / 2;
  np = (np + kPairUnroll - 1) / kPairUnroll * kPairUnroll;
  if (np == 0) np = kPairUnroll;
  p.npairs = np;
  p.pairs.assign((size_t)(np + kPairPad) * 8, 0.0f);
  const double u = 5.9604644775390625e-8;   // 2^-24
  for (uint32_t i = 0; i < (np + kPairPad) * 2; ++i) {
    float ncx = 0.f, ncy = 0.f, ncz = 0.f, nk = -1e30f;   // padding: cc = +1e30 -> never a candidate
    if (i < sc.n) {
      const double cx = sc.geom[i].x, cy = sc.geom[i].y, cz = sc.geom[i].z, r = sc.geom[i].w;
      const double c2 = cx * cx + cy * cy + cz * cz;
      const double k = c2 - r * r;
      const double bound = 88.0 * u * c2 + 10.0 * u * r * r;
      ncx = -sc.geom[i].x; ncy = -sc.geom[i].y; ncz = -sc.geom[i].z;
      nk = round_up_to_float(-k + bound);
    }
    const uint32_t pair = i / 2, half = i & 1;
    float* q = &p.pairs[(size_t)pair * 8];
    q[0 + half] = ncx; q[2 + half] = ncy; q[4 + half] = ncz; q[6 + half] = nk;
  }
  return p;
}

}  // namespace uecrt
