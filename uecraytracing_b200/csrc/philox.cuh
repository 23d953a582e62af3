// philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., "Parallel random numbers: as
// easy as 1, 2, 3", SC'11), replacing the reference's per-sample yk::mt19937 construction
// (source.cpp:154-160, random.hpp:43-151: 624-word seed + twist per (pixel,sample), ~80% of
// the reference's CPU time).
//
// Stream layout used by the path tracer (north_star: "keyed by (pixel, sample, bounce)"):
//   key     = {seed_lo, seed_hi}
//   counter = {pixel, sample, bounce, sub}
//     bounce 0      : camera  -> {jitter_u, jitter_v, lens_radius^2, lens_angle/2pi}
//     bounce b >= 1 : scatter at the b-th hit -> lambertian x,y,z | fuzz x,y,z,radius | dielectric xi
// One 128-bit block per event, so a path never shares random numbers with any other
// (pixel, sample) and the image is independent of how samples are sharded over GPUs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define UECRT_HD __host__ __device__ __forceinline__
#else
#define UECRT_HD inline
#endif

struct philox_out { uint32_t x, y, z, w; };

// 32x32 -> 64-bit product: one IMAD.WIDE.U32 on the device (hi and lo halves of a Philox round)
UECRT_HD uint64_t uecrt_mulwide32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint64_t p;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
  return p;
#else
  return (uint64_t)a * (uint64_t)b;
#endif
}

UECRT_HD philox_out philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = uecrt_mulwide32(M0, c0), p1 = uecrt_mulwide32(M1, c2);
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  return philox_out{c0, c1, c2, c3};
}

// 24-bit uniform in [0,1): exactly representable, never 1.0
UECRT_HD float uecrt_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
