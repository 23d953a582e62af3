// closest_hit.cuh -- device closest-hit search over the flattened sphere list.
//
// Replaces hittable_list::hit_impl (yk/hittable_list.hpp:32-58: left fold in insertion order
// with a shrinking CLOSED interval, later index wins exact ties) over sphere::hit_impl
// (yk/sphere.hpp:25-48).  Two phases per ray:
//
//  1. SCAN (the roofline loop, FP32, packed f32x2 -> SASS FFMA2/FADD2): for every sphere PAIR
//     read two broadcast LDS.128 from shared memory and evaluate, with the ray direction
//     normalised (a = 1),
//         hb   = o.d - c.d                      (3 FFMA2, chain starts at the per-ray o.d)
//         -cc  = -(|o|^2 - 2 o.c + |c|^2 - r^2) (3 FFMA2 + 1 FADD2; |c|^2-r^2 is per-sphere)
//         disc = hb^2 - cc                      (1 FFMA2)
//     i.e. 8 packed FP32 instructions per 2 tests.  The per-sphere constant is stored
//     RAISED and |o|^2 LOWERED by a rigorous bound on the FP32 rounding error of this
//     expression (scene_pack.hpp), so "disc_fp32 >= 0" is a SUPERSET of the reference's
//     "discriminant >= 0" (sphere.hpp:32).  Sign bits of the two discriminants are combined
//     with one LOP3; a taken branch only appends the sphere index to a per-thread candidate
//     list in shared memory.  Nothing else happens in the loop: every non-FP32 instruction
//     costs an FP32 issue slot on sm_100 (tools/microbench, profiles/ubench_r1.log).
//  2. RESOLVE (once per ray, a few candidates): walk the candidate list in index order and
//     apply the reference's root selection and closed-interval update
//       FAST : FP32, cancellation-free form  l = oc - (oc.d)d, disc = r^2 - |l|^2
//       EXACT: FP64 with the reference's exact operation order (no FMA contraction), so the
//              winning index equals the double-precision reference's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uecrt {

constexpr int kListCap = 16;   // candidate-list entries per thread (uint16 each, in shared memory)

// ---- packed f32x2 helpers (PTX ISA 8.6+, sm_100: SASS FFMA2 / FADD2) -------------------
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Per-ray constants of the scan (unit direction).
struct ScanRay {
  float dx, dy, dz;     // unit direction
  float m2ox, m2oy, m2oz;  // -2 * origin
  float od;             // origin . direction
  float noo;            // -(|origin|^2 * (1 - kRayBound))   (lowered |o|^2, negated)
};

// Fraction of |o|^2 by which the ray part of the discriminant bound lowers cc (scene_pack.hpp).
constexpr float kRayBound = 80.0f * 5.9604644775390625e-8f;   // 80 * 2^-24

__device__ __forceinline__ ScanRay make_scan_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
  ScanRay s;
  s.dx = dx; s.dy = dy; s.dz = dz;
  s.m2ox = -2.0f * ox; s.m2oy = -2.0f * oy; s.m2oz = -2.0f * oz;
  s.od = fmaf(oz, dz, fmaf(oy, dy, ox * dx));
  const float oo = fmaf(oz, oz, fmaf(oy, oy, ox * ox));
  s.noo = -(oo - oo * kRayBound);
  return s;
}

// Candidate list: entry e of thread t lives at list[e * stride] (list already offset by t).
struct CandList {
  uint16_t* base;
  int stride;   // in uint16 units (= threads per block)
};

// ---------------------------------------------------------------------------------------
// Phase 2, FAST: FP32 resolve of `cnt` candidates (ascending index order).
// t is the distance along the UNIT direction; tmin is 0.001*|d| in those units.
struct FastBest { float t; int id; };

__device__ __forceinline__ void resolve_fast(const float4* __restrict__ geom, const CandList& L, int cnt,
                                             float ox, float oy, float oz, float dx, float dy, float dz,
                                             float tmin, int self_id, FastBest& best) {
  for (int e = 0; e < cnt; ++e) {
    const int j = L.base[e * L.stride];
    const float4 g = __ldg(&geom[j]);
    const float ocx = ox - g.x, ocy = oy - g.y, ocz = oz - g.z;
    const float hb = fmaf(ocz, dz, fmaf(ocy, dy, ocx * dx));
    float t;
    if (j == self_id) {
      // the ray starts ON this sphere: the roots are exactly 0 and -2*hb (reference: the
      // ~1e-16 self-hit is below t_min = 0.001 and only the far root can be accepted)
      t = -2.0f * hb;
      if (!(t >= tmin)) continue;
    } else {
      const float lx = fmaf(-hb, dx, ocx), ly = fmaf(-hb, dy, ocy), lz = fmaf(-hb, dz, ocz);
      const float disc = fmaf(g.w, g.w, -fmaf(lz, lz, fmaf(ly, ly, lx * lx)));
      if (disc < 0.0f) continue;                       // sphere.hpp:32
      const float s = sqrtf(disc);
      t = -hb - s;                                      // sphere.hpp:35
      if (t < tmin) {                                   // sphere.hpp:36-39 (t_max test below)
        t = -hb + s;
        if (t < tmin) continue;
      }
    }
    if (t <= best.t) { best.t = t; best.id = j; }       // closed interval: later index wins ties
  }
}

// ---------------------------------------------------------------------------------------
// Phase 2, EXACT: FP64 with the reference's operation order.  (dx,dy,dz) is the direction the
// caller wants t measured in (the ORIGINAL unnormalised direction for K2, so t == the
// reference's root; the unit direction inside K1).
struct ExactBest { double t; int id; };

__device__ __forceinline__ void resolve_exact(const float4* __restrict__ geom, const CandList& L, int cnt,
                                              double ox, double oy, double oz, double dx, double dy, double dz,
                                              double a, double tmin, ExactBest& best) {
  for (int e = 0; e < cnt; ++e) {
    const int j = L.base[e * L.stride];
    const float4 g = __ldg(&geom[j]);
    const double cx = g.x, cy = g.y, cz = g.z, r = g.w;
    const double ocx = __dsub_rn(ox, cx), ocy = __dsub_rn(oy, cy), ocz = __dsub_rn(oz, cz);     // sphere.hpp:27
    const double half_b = __dadd_rn(__dadd_rn(__dmul_rn(ocx, dx), __dmul_rn(ocy, dy)), __dmul_rn(ocz, dz));  // :29
    const double oc2 = __dadd_rn(__dadd_rn(__dmul_rn(ocx, ocx), __dmul_rn(ocy, ocy)), __dmul_rn(ocz, ocz));
    const double c = __dsub_rn(oc2, __dmul_rn(r, r));                                            // :30
    const double disc = __dsub_rn(__dmul_rn(half_b, half_b), __dmul_rn(a, c));                   // :31
    if (disc < 0.0) continue;                                                                     // :32
    const double sqrtd = __dsqrt_rn(disc);                                                        // :33
    double root = __ddiv_rn(__dsub_rn(-half_b, sqrtd), a);                                       // :35
    if (root < tmin || best.t < root) {                                                           // :36
      root = __ddiv_rn(__dadd_rn(-half_b, sqrtd), a);                                            // :37
      if (root < tmin || best.t < root) continue;                                                 // :38
    }
    best.t = root;                                                                                // hittable_list.hpp:40-41
    best.id = j;
  }
}

// ---------------------------------------------------------------------------------------
// Phase 1: the scan.  pairs: shared memory, 2 x 16 B per sphere pair
//   A = {-cx0, -cx1, -cy0, -cy1},  B = {-cz0, -cz1, nk0, nk1},  nk = -(|c|^2 - r^2 - bound)
// npairs is a multiple of 4 (host pads with never-hit pairs).  `index_base` is added to
// stored indices (tile offset when the scene is swept in tiles).
// On list overflow the FULL callback folds the list into the running best and empties it.
template <class OnFull>
__device__ __forceinline__ int scan_pairs(const ulonglong2* __restrict__ pairs, int npairs, int index_base,
                                          const ScanRay& r, const CandList& L, int cnt, OnFull&& on_full) {
  const unsigned long long dx2 = pk2(r.dx, r.dx), dy2 = pk2(r.dy, r.dy), dz2 = pk2(r.dz, r.dz);
  const unsigned long long ox2 = pk2(r.m2ox, r.m2ox), oy2 = pk2(r.m2oy, r.m2oy), oz2 = pk2(r.m2oz, r.m2oz);
  const unsigned long long od2 = pk2(r.od, r.od), noo2 = pk2(r.noo, r.noo);
#pragma unroll 4
  for (int j = 0; j < npairs; ++j) {
    const ulonglong2 A = pairs[2 * j];
    const ulonglong2 B = pairs[2 * j + 1];
    unsigned long long hb = fma2(A.x, dx2, od2);
    unsigned long long ncc = fma2(A.x, ox2, noo2);
    hb = fma2(A.y, dy2, hb);
    ncc = fma2(A.y, oy2, ncc);
    hb = fma2(B.x, dz2, hb);
    ncc = fma2(B.x, oz2, ncc);
    ncc = add2(ncc, B.y);
    const unsigned long long disc = fma2(hb, hb, ncc);
    const uint32_t d0 = (uint32_t)disc, d1 = (uint32_t)(disc >> 32);
    if ((int)(d0 & d1) >= 0) {              // at least one sign bit clear: a candidate
      if ((int)d0 >= 0) {
        if (cnt == kListCap) { on_full(cnt); cnt = 0; }
        L.base[cnt * L.stride] = (uint16_t)(index_base + 2 * j);
        ++cnt;
      }
      if ((int)d1 >= 0) {
        if (cnt == kListCap) { on_full(cnt); cnt = 0; }
        L.base[cnt * L.stride] = (uint16_t)(index_base + 2 * j + 1);
        ++cnt;
      }
    }
  }
  return cnt;
}

// Hit record in FP32 (yk/hittable.hpp:16-28; id is the OBJECT index as in hittable_list.hpp:41)
struct Hit {
  float t;            // distance along the unit direction
  float px, py, pz;
  float nx, ny, nz;   // unit normal facing against the ray
  int id;             // -1 = miss
  bool front;
};

// sphere.hpp:41-45 + hittable.hpp:23-27 for the winning sphere, normal taken from the
// centre-relative vector (never from the rounded world-space hit point).
__device__ __forceinline__ void finish_hit_fast(const float4* __restrict__ geom, float ox, float oy, float oz,
                                                float dx, float dy, float dz, Hit& h) {
  const float4 g = __ldg(&geom[h.id]);
  const float ocx = ox - g.x, ocy = oy - g.y, ocz = oz - g.z;
  const float inv_r = __fdividef(1.0f, g.w);
  h.px = fmaf(h.t, dx, ox); h.py = fmaf(h.t, dy, oy); h.pz = fmaf(h.t, dz, oz);
  float nx = fmaf(h.t, dx, ocx) * inv_r, ny = fmaf(h.t, dy, ocy) * inv_r, nz = fmaf(h.t, dz, ocz) * inv_r;
  // |n| is 1 up to the rounding of t; renormalise so scatter sees a unit normal
  const float s = rsqrtf(fmaf(nz, nz, fmaf(ny, ny, nx * nx)));
  nx *= s; ny *= s; nz *= s;
  h.front = fmaf(dz, nz, fmaf(dy, ny, dx * nx)) < 0.0f;
  h.nx = h.front ? nx : -nx; h.ny = h.front ? ny : -ny; h.nz = h.front ? nz : -nz;
}

}  // namespace uecrt
