// closest_hit.cuh -- device closest-hit search over the flattened sphere list.
//
// Replaces hittable_list::hit_impl (yk/hittable_list.hpp:32-58: left fold in insertion order
// with a shrinking CLOSED interval, later index wins exact ties) over sphere::hit_impl
// (yk/sphere.hpp:25-48).  Two phases per ray:
//
//  1. SCAN (the roofline loop, FP32, packed f32x2 -> SASS FMUL2/FFMA2): with the ray direction
//     normalised (a = 1) the reference discriminant half_b^2 - a*c (sphere.hpp:27-31) is
//         disc = (c.d)^2 + q.c - k - P,   q = 2(o - (o.d)d),  k = |c|^2 - r^2,  P = |o|^2 - (o.d)^2
//     q and P depend on the ray only, k on the sphere only, so per sphere PAIR the loop reads
//     two broadcast LDS.128 from shared memory and evaluates
//         cd = c.d                 (FMUL2 + 2 FFMA2)
//         e  = q.c - k             (3 FFMA2, the chain starts at the per-sphere -k)
//         f  = cd^2 + e            (1 FFMA2)
//     i.e. 7 packed FP32 instructions per 2 tests, and a sphere is a candidate iff f >= P.
//     -k is stored RAISED and P LOWERED by a rigorous bound on the FP32 rounding error of this
//     expression (scene_pack.hpp), so {f >= P} is a SUPERSET of the reference's
//     "discriminant >= 0" (sphere.hpp:32).  P > 0, so f >= P is an INTEGER compare of the float
//     bit patterns: eight f are reduced with 3-input integer max and one branch per 8 spheres
//     guards the code that appends sphere indices to a per-thread candidate list in shared
//     memory.  Nothing else happens in the loop: every non-FP32 instruction costs an FP32 issue
//     slot on sm_100 (tools/microbench, profiles/ubench_r1.log).
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
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 1/sqrt(x): MUFU.RSQ plus one Newton step (~1 ulp), instead of the multi-instruction IEEE
// sqrtf + divide; the scan's error bound budgets 3u for the normalised direction.
__device__ __forceinline__ float rsqrt_nr(float x) {
  const float y = rsqrtf(x);
  return y * fmaf(-0.5f * x, y * y, 1.5f);
}
// sqrt(x) for x >= 0 through MUFU.SQRT (sqrt.approx, <= 2 ulp); exact 0 stays 0
__device__ __forceinline__ float sqrt_fast(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Per-ray constants of the scan (unit direction).
struct ScanRay {
  float dx, dy, dz;        // unit direction
  float nqx, nqy, nqz;     // -q = 2((o.d)d - o)
  int thr;                 // bit pattern of P lowered by the ray part of the error bound, clamped to >= +0
  float p_low;             // the same lowered P before the clamp (tensor-core scan: subtracted inside the MMA)
};

// Fraction of |o|^2 by which the ray part of the discriminant bound lowers P (scene_pack.hpp: the FP32 scan needs 62u,
// the tensor scan 204u; both use this constant).
constexpr float kRayBound = 208.0f * 5.9604644775390625e-8f;   // 208 * 2^-24

__device__ __forceinline__ ScanRay make_scan_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
  ScanRay s;
  s.dx = dx; s.dy = dy; s.dz = dz;
  const float od = fmaf(oz, dz, fmaf(oy, dy, ox * dx));
  s.nqx = 2.0f * fmaf(od, dx, -ox); s.nqy = 2.0f * fmaf(od, dy, -oy); s.nqz = 2.0f * fmaf(od, dz, -oz);
  const float oo = fmaf(oz, oz, fmaf(oy, oy, ox * ox));
  const float P = fmaf(-od, od, oo) - oo * kRayBound;
  // f >= P as a signed-integer compare needs P >= +0: a lowered P below zero becomes +0, which only
  // adds candidates (every f with a clear sign bit), never drops one
  s.thr = __float_as_int(fmaxf(P, 0.0f));
  s.p_low = P;
  return s;
}

// Candidate list: entry e of thread t lives at base[e * STRIDE] (base already offset by t);
// STRIDE = threads per block, so the 32 lanes of a warp write 32 consecutive uint16.
template <int STRIDE>
struct CandList {
  uint16_t* base;
  static constexpr int stride = STRIDE;
};

// ---------------------------------------------------------------------------------------
// Phase 2, FAST: FP32 resolve of `cnt` candidates (ascending index order).
// t is the distance along the UNIT direction; tmin is 0.001*|d| in those units.
struct FastBest { float t; int id; bool far_root = false; };   // far_root: the accepted root is -hb + s (refine_winner solves the same one)

// one candidate sphere j against the running best (sphere.hpp:32-39 + hittable_list.hpp:37-44).
// ORDERED = true: candidates arrive in ascending index order, so accepting t <= best is the reference's
// "later index wins"; ORDERED = false (grid walk: arbitrary order, a sphere may come twice): the same rule
// spelled out, t < best or (t == best and j > best.id).
template <bool ORDERED = true>
__device__ __forceinline__ void test_fast_g(const float4 g, int j, float ox, float oy, float oz, float dx,
                                            float dy, float dz, float tmin, int self_id, FastBest& best) {
  const float ocx = ox - g.x, ocy = oy - g.y, ocz = oz - g.z;
  const float hb = fmaf(ocz, dz, fmaf(ocy, dy, ocx * dx));
  float t;
  bool far_root = true;
  if (j == self_id) {
    // the ray starts ON this sphere: the roots are exactly 0 and -2*hb (reference: the
    // ~1e-16 self-hit is below t_min = 0.001 and only the far root can be accepted)
    t = -2.0f * hb;
    if (!(t >= tmin)) return;
  } else {
    const float lx = fmaf(-hb, dx, ocx), ly = fmaf(-hb, dy, ocy), lz = fmaf(-hb, dz, ocz);
    const float disc = fmaf(g.w, g.w, -fmaf(lz, lz, fmaf(ly, ly, lx * lx)));
    if (disc < 0.0f) return;                         // sphere.hpp:32
    const float s = sqrt_fast(disc);
    t = -hb - s;                                      // sphere.hpp:35
    far_root = t < tmin;
    if (far_root) {                                   // sphere.hpp:36-39 (t_max test below)
      t = -hb + s;
      if (t < tmin) return;
    }
  }
  if (ORDERED ? (t <= best.t) : (t < best.t || (t == best.t && j > best.id))) { best.t = t; best.id = j; best.far_root = far_root; }   // closed interval: later index wins ties
}
template <bool ORDERED = true>
__device__ __forceinline__ void test_fast(const float4* __restrict__ geom, int j, float ox, float oy, float oz, float dx,
                                          float dy, float dz, float tmin, int self_id, FastBest& best) {
  test_fast_g<ORDERED>(__ldg(&geom[j]), j, ox, oy, oz, dx, dy, dz, tmin, self_id, best);
}

template <int STRIDE>
__device__ __forceinline__ void resolve_fast(const float4* __restrict__ geom, const CandList<STRIDE>& L, int cnt,
                                             float ox, float oy, float oz, float dx, float dy, float dz,
                                             float tmin, int self_id, FastBest& best) {
  for (int e = 0; e < cnt; ++e) test_fast(geom, L.base[e * STRIDE], ox, oy, oz, dx, dy, dz, tmin, self_id, best);
}
// ---------------------------------------------------------------------------------------
// Phase 2, EXACT: FP64 with the reference's operation order.  (dx,dy,dz) is the direction the
// caller wants t measured in (the ORIGINAL unnormalised direction for K2, so t == the
// reference's root; the unit direction inside K1).
struct ExactBest { double t; int id; };

// self_id (K1 only): the sphere the ray starts ON.  The origin is the hit point rounded to FP32, so its near root on that
// sphere is not the reference's ~1e-16 but up to ~1e-3 at grazing angles on a large sphere (r = 1000 ground): above t_min,
// a false self-hit the double reference cannot have.  As in the FAST test, only the far root of that sphere may be accepted.
template <bool ORDERED = true>
__device__ __forceinline__ void test_exact_g(const float4 g, int j, double ox, double oy, double oz, double dx,
                                             double dy, double dz, double a, double tmin, ExactBest& best, int self_id = -1) {
  const double cx = g.x, cy = g.y, cz = g.z, r = g.w;
  const double ocx = __dsub_rn(ox, cx), ocy = __dsub_rn(oy, cy), ocz = __dsub_rn(oz, cz);     // sphere.hpp:27
  const double half_b = __dadd_rn(__dadd_rn(__dmul_rn(ocx, dx), __dmul_rn(ocy, dy)), __dmul_rn(ocz, dz));  // :29
  const double oc2 = __dadd_rn(__dadd_rn(__dmul_rn(ocx, ocx), __dmul_rn(ocy, ocy)), __dmul_rn(ocz, ocz));
  const double c = __dsub_rn(oc2, __dmul_rn(r, r));                                            // :30
  const double disc = __dsub_rn(__dmul_rn(half_b, half_b), __dmul_rn(a, c));                   // :31
  if (disc < 0.0) return;                                                                       // :32
  const double sqrtd = __dsqrt_rn(disc);                                                        // :33
  double root = __ddiv_rn(__dsub_rn(-half_b, sqrtd), a);                                       // :35
  if (j == self_id || root < tmin || best.t < root) {                                           // :36
    root = __ddiv_rn(__dadd_rn(-half_b, sqrtd), a);                                            // :37
    if (root < tmin || best.t < root) return;                                                   // :38
  }
  if (!ORDERED && root == best.t && j <= best.id) return;   // unordered candidates: an equal root only wins with a LATER index
  best.t = root;                                                                                // hittable_list.hpp:40-41
  best.id = j;
}
template <bool ORDERED = true>
__device__ __forceinline__ void test_exact(const float4* __restrict__ geom, int j, double ox, double oy, double oz, double dx,
                                           double dy, double dz, double a, double tmin, ExactBest& best, int self_id = -1) {
  test_exact_g<ORDERED>(__ldg(&geom[j]), j, ox, oy, oz, dx, dy, dz, a, tmin, best, self_id);
}

template <int STRIDE>
__device__ __forceinline__ void resolve_exact(const float4* __restrict__ geom, const CandList<STRIDE>& L, int cnt,
                                              double ox, double oy, double oz, double dx, double dy, double dz,
                                              double a, double tmin, ExactBest& best, int self_id = -1) {
  for (int e = 0; e < cnt; ++e) test_exact(geom, L.base[e * STRIDE], ox, oy, oz, dx, dy, dz, a, tmin, best, self_id);
}
// ---------------------------------------------------------------------------------------
// Phase 1: the scan.  pairs: shared memory, 2 x 16 B per sphere pair
//   A = {-cx0, -cx1, -cy0, -cy1},  B = {-cz0, -cz1, nk0, nk1},  nk = -(|c|^2 - r^2) + bound
// Two pairs (4 spheres) form a GROUP; npairs is a multiple of 8 and the host appends one more
// never-hit group so the loop can prefetch group g+1 unconditionally.  A thread scans R rays at
// once (R = 1 or 2): every sphere load, the loop control and the branch are shared by the R rays,
// which is what amortises the non-FP32 instructions of the loop (each costs an FP32 issue slot).
// Per 8 spheres the loop issues 8 LDS.128 for the NEXT iteration (software pipelining: the loads
// are in flight across this iteration's math), R x 28 packed FP32 ops in 2R interleaved dependency
// chains, R x 5 integer max, R compares and one branch.  `index_base` is added to stored indices
// (tile offset when the scene is swept in tiles); cnt[r] carries the candidates already in ray r's
// list.  When a list is nearly full the FULL callback folds it into that ray's running best and
// the list restarts.
template <int R, int STRIDE, class OnFull>
__device__ __forceinline__ void scan_pairs(const ulonglong2* __restrict__ pairs, int npairs, int index_base,
                                           const ScanRay (&ray)[R], const CandList<STRIDE> (&L)[R], int (&cnt)[R],
                                           OnFull&& on_full) {
  unsigned long long dx2[R], dy2[R], dz2[R], qx2[R], qy2[R], qz2[R];
  int thr[R];
  // byte offset of the next free list entry (32-bit shared-memory arithmetic only)
  constexpr uint32_t kEntryBytes = (uint32_t)STRIDE * (uint32_t)sizeof(uint16_t);
  constexpr uint32_t kFullAt = (uint32_t)(kListCap - 8) * kEntryBytes;   // < 8 free entries left
  unsigned char* lbase[R];
  uint32_t ofs[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    dx2[r] = pk2(ray[r].dx, ray[r].dx); dy2[r] = pk2(ray[r].dy, ray[r].dy); dz2[r] = pk2(ray[r].dz, ray[r].dz);
    qx2[r] = pk2(ray[r].nqx, ray[r].nqx); qy2[r] = pk2(ray[r].nqy, ray[r].nqy); qz2[r] = pk2(ray[r].nqz, ray[r].nqz);
    thr[r] = ray[r].thr;
    lbase[r] = reinterpret_cast<unsigned char*>(L[r].base);
    ofs[r] = (uint32_t)cnt[r] * kEntryBytes;   // candidates carried over from earlier tiles
  }

  // f for the four spheres of one group and ray r: two interleaved chains of 7 packed ops
#define UECRT_GROUP_MATH(r, A0, B0, A1, B1, f0, f1, f2, f3)                                \
  {                                                                                        \
    unsigned long long cd0 = mul2(A0.x, dx2[r]);                                           \
    unsigned long long cd1 = mul2(A1.x, dx2[r]);                                           \
    unsigned long long e0 = fma2(A0.x, qx2[r], B0.y);                                      \
    unsigned long long e1 = fma2(A1.x, qx2[r], B1.y);                                      \
    cd0 = fma2(A0.y, dy2[r], cd0);                                                         \
    cd1 = fma2(A1.y, dy2[r], cd1);                                                         \
    e0 = fma2(A0.y, qy2[r], e0);                                                           \
    e1 = fma2(A1.y, qy2[r], e1);                                                           \
    cd0 = fma2(B0.x, dz2[r], cd0);                                                         \
    cd1 = fma2(B1.x, dz2[r], cd1);                                                         \
    e0 = fma2(B0.x, qz2[r], e0);                                                           \
    e1 = fma2(B1.x, qz2[r], e1);                                                           \
    const unsigned long long g0 = fma2(cd0, cd0, e0);                                      \
    const unsigned long long g1 = fma2(cd1, cd1, e1);                                      \
    asm("mov.b64 {%0, %1}, %2;" : "=r"(f0), "=r"(f1) : "l"(g0));                           \
    asm("mov.b64 {%0, %1}, %2;" : "=r"(f2), "=r"(f3) : "l"(g1));                           \
  }
#define UECRT_PUSH(r, f, idx)                                                              \
  if (f >= thr[r]) { *reinterpret_cast<uint16_t*>(lbase[r] + ofs[r]) = (uint16_t)(idx); ofs[r] += kEntryBytes; }

  const ulonglong2* p = pairs;
  ulonglong2 A0 = p[0], B0 = p[1], A1 = p[2], B1 = p[3];
  const int niter = npairs >> 2;          // 8 spheres per iteration; npairs is a multiple of 8
#ifndef UECRT_SCAN_UNROLL
#define UECRT_SCAN_UNROLL 2               // tuning knob (1, 2: npairs is padded to 8 pairs)
#endif
#define UECRT_PRAGMA_(x) _Pragma(#x)
#define UECRT_UNROLL_(n) UECRT_PRAGMA_(unroll n)
  UECRT_UNROLL_(UECRT_SCAN_UNROLL)
  for (int it = 0; it < niter; ++it) {
    int f[R][8], mlo[R], mhi[R];
    const ulonglong2 A2 = p[4], B2 = p[5], A3 = p[6], B3 = p[7];          // second group of this iteration
#pragma unroll
    for (int r = 0; r < R; ++r) UECRT_GROUP_MATH(r, A0, B0, A1, B1, f[r][0], f[r][1], f[r][2], f[r][3])
    p += 8;
    A0 = p[0]; B0 = p[1]; A1 = p[2]; B1 = p[3];                            // first group of the next (padded)
#pragma unroll
    for (int r = 0; r < R; ++r) UECRT_GROUP_MATH(r, A2, B2, A3, B3, f[r][4], f[r][5], f[r][6], f[r][7])
    bool any = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mlo[r] = max(max(f[r][0], f[r][1]), max(f[r][2], f[r][3]));
      mhi[r] = max(max(f[r][4], f[r][5]), max(f[r][6], f[r][7]));
      any = any || (max(mlo[r], mhi[r]) >= thr[r]);
    }
    if (any) {                            // at least one of the 8 spheres is a candidate for one of the rays
      const int i0 = index_base + 8 * it;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (R > 1 && max(mlo[r], mhi[r]) < thr[r]) continue;
        if (ofs[r] > kFullAt) { on_full(r, (int)(ofs[r] / kEntryBytes)); ofs[r] = 0; }
        if (mlo[r] >= thr[r]) {
          UECRT_PUSH(r, f[r][0], i0) UECRT_PUSH(r, f[r][1], i0 + 1) UECRT_PUSH(r, f[r][2], i0 + 2) UECRT_PUSH(r, f[r][3], i0 + 3)
        }
        if (mhi[r] >= thr[r]) {
          UECRT_PUSH(r, f[r][4], i0 + 4) UECRT_PUSH(r, f[r][5], i0 + 5) UECRT_PUSH(r, f[r][6], i0 + 6) UECRT_PUSH(r, f[r][7], i0 + 7)
        }
      }
    }
  }
#undef UECRT_GROUP_MATH
#undef UECRT_PUSH
#pragma unroll
  for (int r = 0; r < R; ++r) cnt[r] = (int)(ofs[r] / kEntryBytes);
}

// Hit record in FP32 (yk/hittable.hpp:16-28; id is the OBJECT index as in hittable_list.hpp:41)
struct Hit {
  float t;            // distance along the unit direction
  float px, py, pz;
  float nx, ny, nz;   // unit normal facing against the ray
  int id;             // -1 = miss
  bool front;
};

// The WINNER's root and normal in double (sphere.hpp:27-45 for one sphere).  The candidate test (test_fast_g) computes
// t = -hb -+ sqrt(disc) in FP32, which loses digits by cancellation when the sphere is much larger than the distance to
// it (the r = 1000 ground seen from 10 units: 5e-4 relative on the final scene's primary rays, 5e-3 on horizon rays).
// Once per ray, for the sphere that won, the quadratic is solved again in FP64 for the direction vector (vx, vy, vz) the
// caller wants t measured in (K1: its FP32 unit direction; K2: the caller's original direction) and the SAME root (near
// or far, FastBest::far_root) as the FP32 test took is returned: t and the normal are then exact for the given ray up to double rounding.  Which sphere wins is
// still decided by the FP32 roots, identically in every scan mode.  RECIP: a = |v|^2 is 1 +- 1e-7, so 1 / a = 2 - a to
// 1e-14 (K1); otherwise a true division (K2).
struct HitD { double t, nx, ny, nz; bool front; };
template <bool RECIP>
__device__ __forceinline__ HitD refine_winner(const float4 g, float ox, float oy, float oz, float vx, float vy, float vz, bool near_root) {
  const double Dx = vx, Dy = vy, Dz = vz, r = g.w;
  const double ocx = (double)ox - (double)g.x, ocy = (double)oy - (double)g.y, ocz = (double)oz - (double)g.z;      // sphere.hpp:27
  const double a = fma(Dz, Dz, fma(Dy, Dy, Dx * Dx));                                                             // :28
  const double hb = fma(ocz, Dz, fma(ocy, Dy, ocx * Dx));                                                         // :29
  const double c = fma(ocz, ocz, fma(ocy, ocy, fma(ocx, ocx, -r * r)));                                           // :30
  const double disc = fma(hb, hb, -a * c);                                                                        // :31
  const double sq = sqrt(fmax(disc, 0.0));                                                                        // :33
  double r0, r1;
  if (RECIP) { const double ia = 2.0 - a; r0 = (-hb - sq) * ia; r1 = (-hb + sq) * ia; }
  else { r0 = (-hb - sq) / a; r1 = (-hb + sq) / a; }                                                              // :35, :37
  HitD h;
  h.t = near_root ? r0 : r1;
  const double ir = RECIP ? (double)__fdividef(1.0f, g.w) : 1.0 / r;
  h.nx = fma(h.t, Dx, ocx) * ir; h.ny = fma(h.t, Dy, ocy) * ir; h.nz = fma(h.t, Dz, ocz) * ir;                     // :44 (centre-relative)
  h.front = fma(Dz, h.nz, fma(Dy, h.ny, Dx * h.nx)) < 0.0;                                                        // hittable.hpp:25
  if (!h.front) { h.nx = -h.nx; h.ny = -h.ny; h.nz = -h.nz; }
  return h;
}

// K1: hit record of the winning sphere for the FP32 unit direction (dx, dy, dz); far_root: which root the FP32 test took.
// Only the root is solved in double (refine_winner's :27-:37 with |d|^2 taken as 1: K1's directions are normalised in
// FP32, |d|^2 = 1 +- 1.2e-7, which bounds what that costs; the square root is one FP64 Newton step on MUFU.RSQ, 46 bits);
// the point and the normal follow in FP32 from the centre-relative vector, exact to 1e-7 once t is.
__device__ __forceinline__ void finish_hit_fast_g(const float4 g, float ox, float oy, float oz,
                                                  float dx, float dy, float dz, bool far_root, Hit& h) {
  const float ocx = ox - g.x, ocy = oy - g.y, ocz = oz - g.z;
  {
    const double Dx = dx, Dy = dy, Dz = dz, r = g.w;
    const double Ocx = (double)ox - (double)g.x, Ocy = (double)oy - (double)g.y, Ocz = (double)oz - (double)g.z;
    const double hb = fma(Ocz, Dz, fma(Ocy, Dy, Ocx * Dx));
    const double c = fma(Ocz, Ocz, fma(Ocy, Ocy, fma(Ocx, Ocx, -r * r)));
    const double disc = fmax(fma(hb, hb, -c), 0.0);
    double y = (double)rsqrtf(fmaxf((float)disc, 1e-37f));                      // 1 / sqrt(disc) to 2^-23
    y = fma(0.5 * y, fma(-disc * y, y, 1.0), y);                                // Newton: 2^-45
    const double sq = disc * y;
    h.t = (float)(far_root ? sq - hb : -hb - sq);
  }
  const float inv_r = __fdividef(1.0f, g.w);
  h.px = fmaf(h.t, dx, ox); h.py = fmaf(h.t, dy, oy); h.pz = fmaf(h.t, dz, oz);
  float nx = fmaf(h.t, dx, ocx) * inv_r, ny = fmaf(h.t, dy, ocy) * inv_r, nz = fmaf(h.t, dz, ocz) * inv_r;
  // |n| is 1 up to the rounding of t; renormalise so scatter sees a unit normal
  const float s = rsqrtf(fmaf(nz, nz, fmaf(ny, ny, nx * nx)));
  nx *= s; ny *= s; nz *= s;
  h.front = fmaf(dz, nz, fmaf(dy, ny, dx * nx)) < 0.0f;
  h.nx = h.front ? nx : -nx; h.ny = h.front ? ny : -ny; h.nz = h.front ? nz : -nz;
}
__device__ __forceinline__ void finish_hit_fast(const float4* __restrict__ geom, float ox, float oy, float oz,
                                                float dx, float dy, float dz, bool far_root, Hit& h) {
  finish_hit_fast_g(__ldg(&geom[h.id]), ox, oy, oz, dx, dy, dz, far_root, h);
}

}  // namespace uecrt
