// tc_collect.cuh -- consumer side of the tensor-core scan: one ray per thread, candidates into the
// same per-thread shared-memory list the FP32 scan fills (closest_hit.cuh), so the resolve phase
// (reference root selection, closed interval, later index wins: hittable_list.hpp:32-58,
// sphere.hpp:32-39) is shared by both scans.
#pragma once
#include "closest_hit.cuh"
#include "tc_scan.cuh"

namespace uecrt {
namespace tc {

// One round: publish this thread's lifted ray as row `c.row` of the group's A tile, then read the
// f values of this ray against every sphere chunk from TMEM as the issuer's MMAs complete.
// EVERY consumer thread of the group must call this once per round (a thread without a live ray
// passes ray.thr = +inf).  cnt returns the number of candidates left in the list.
template <int STRIDE, class OnFull>
__device__ __forceinline__ void collect(Consumer& c, int nchunks, const ScanRay& ray, const CandList<STRIDE>& L, int& cnt,
                                        OnFull&& on_full) {
  {
    const float a[kTerms] = {ray.dx * ray.dx, ray.dy * ray.dy, ray.dz * ray.dz, 2.0f * ray.dx * ray.dy, 2.0f * ray.dx * ray.dz,
                             2.0f * ray.dy * ray.dz, ray.nqx, ray.nqy, ray.nqz, 1.0f};
    float row[kSlots];
    fill_row_a(a, row);
    store_row(c.a_tile, c.row, row);
    fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async proxy
    mbar_arrive(c.a_ready);
  }
  constexpr uint32_t kEntryBytes = (uint32_t)STRIDE * (uint32_t)sizeof(uint16_t);
  constexpr uint32_t kFullAt = (uint32_t)(kListCap - 8) * kEntryBytes;   // < 8 free entries left
  unsigned char* lbase = reinterpret_cast<unsigned char*>(L.base);
  uint32_t ofs = 0;
  const int thr = ray.thr;
  const unsigned lane = threadIdx.x & 31u;
  for (int j = 0; j < nchunks; ++j) {
    const uint32_t s = c.seq++, b = s & 1u;
    mbar_wait(c.full0 + 8u * b, (s >> 1) & 1u);
    fence_after_sync();
#pragma unroll
    for (int blk = 0; blk < kNC / 32; ++blk) {
      int v[32];
      __syncwarp();
      tmem_ld32(c.taddr + b * kNC + blk * 32, v);
      tmem_ld_wait();
      if (blk == kNC / 32 - 1) {       // this warp holds the whole buffer in registers: hand it back
        fence_before_sync();
        if (lane == 0) mbar_arrive(c.empty0 + 8u * b);
      }
      // f >= P as an integer compare of the bit patterns (P >= +0): 8-wide integer maxima, one branch per 32
      int m8[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        m8[k] = max(max(max(v[8 * k], v[8 * k + 1]), max(v[8 * k + 2], v[8 * k + 3])),
                    max(max(v[8 * k + 4], v[8 * k + 5]), max(v[8 * k + 6], v[8 * k + 7])));
      if (max(max(m8[0], m8[1]), max(m8[2], m8[3])) >= thr) {
        const int i0 = j * kNC + blk * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (m8[k] >= thr) {
            if (ofs > kFullAt) { on_full((int)(ofs / kEntryBytes)); ofs = 0; }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (v[8 * k + i] >= thr) { *reinterpret_cast<uint16_t*>(lbase + ofs) = (uint16_t)(i0 + 8 * k + i); ofs += kEntryBytes; }
          }
        }
      }
    }
  }
  cnt = (int)(ofs / kEntryBytes);
}

}  // namespace tc
}  // namespace uecrt
