// tc_collect.cuh -- the SCAN role and the shade-side hand-off of the tensor-core scan.  The scan
// leaves one candidate BITMAP per ray (bit set = the sphere's discriminant may be >= 0); the shade
// thread walks the set bits in ascending sphere index and applies the same per-candidate test as
// the FP32 scan's resolve phase (closest_hit.cuh test_fast / test_exact: reference root selection,
// closed interval, later index wins -- hittable_list.hpp:32-58, sphere.hpp:32-39).
#pragma once
#include "closest_hit.cuh"
#include "tc_scan.cuh"

namespace uecrt {
namespace tc {

// SHADE: publish this thread's ray for the current round: its 32 K slots go to TMEM lane t of the group's
// A tile (warp-collective store).  A thread without a live ray passes live = false: its row is still
// written (every row of the tile takes part in the MMA) with P = +1e30, so all its accumulators
// come out negative and its bitmap is empty.
__device__ __forceinline__ void shade_submit(const Shade& s, const ScanRay& ray, bool live) {
  const float a[kTerms] = {ray.dx * ray.dx, ray.dy * ray.dy, ray.dz * ray.dz, 2.0f * ray.dx * ray.dy, 2.0f * ray.dx * ray.dz,
                           2.0f * ray.dy * ray.dz, ray.nqx, ray.nqy, ray.nqz, 1.0f};
  float row[kSlots];
  fill_row_a(a, live ? ray.p_low : 1e30f, row);
  __syncwarp();
  tmem_st32(s.a_taddr, row);
  tmem_st_wait();
  fence_before_sync();
  mbar_arrive(s.a_ready);
}
// SHADE: wait for the bitmaps of pass m (tc_scan.cuh PassIter: m = the round when the operand is resident)
__device__ __forceinline__ void shade_wait(const Shade& s, uint32_t m) { mbar_wait(s.scan_done, m & 1u); }
// SHADE, swept operand: this thread has consumed its bitmap of the current pass; the ray tile in TMEM stays
// as it is and takes part in the next pass (the arrive releases the bitmap to the scan warps: they only
// write it again behind MMAs the issuer launches after this barrier's phase completes)
__device__ __forceinline__ void shade_next_pass(const Shade& s) { mbar_arrive(s.a_ready); }

// SHADE: f(sphere index) for every candidate of this thread's ray, ascending index.  WARP-CONVERGENT:
// all 32 lanes must call it (a lane without a live ray passes live = false).  Every trip of the
// loop hands one candidate to every lane that still has one (a lane whose current bitmap word is
// exhausted jumps to its next non-empty word through the summary word), so the warp makes
// max-over-lanes(candidates) trips and f runs once per trip with the lanes converged.
template <class F>
__device__ __forceinline__ void for_each_candidate(Shade& s, bool live, F&& f) {
  uint32_t sum = 0u, m = 0u;
  if (live) {
#pragma unroll
    for (int sg = 0; sg < kScanGroups; ++sg) sum |= s.summary[sg * kShadeThreads];
  }
  int w = 0;
  for (;;) {
    if (m == 0u && sum != 0u) {
      w = __clz((int)sum);
      sum &= ~(0x80000000u >> w);
      m = s.bitmap[w * kShadeThreads];
    }
    if (!__any_sync(0xffffffffu, m != 0u)) return;
#ifdef UECRT_TC_PROFILE
    ++s.prof_trips;
    if (m != 0u) ++s.prof_cands;
#endif
    if (m != 0u) {
      const int b = __clz((int)m);
      m &= ~(0x80000000u >> b);
      f(w * 32 + b);
    }
  }
}

// 32 accumulators -> one bitmap word: bit (31 - i) = (v[i] >= 0).  Four independent funnel-shift
// chains (mask = (mask << 1) | sign(v)), merged with byte permutes.
__device__ __forceinline__ uint32_t sign_word(const int (&v)[32]) {
  uint32_t c[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) m = __funnelshift_l((uint32_t)v[8 * k + i], m, 1);
    c[k] = m;
  }
  const uint32_t hi = __byte_perm(c[1], c[0], 0x0040);   // low 16 bits = (c0 << 8) | c1
  const uint32_t lo = __byte_perm(c[3], c[2], 0x0040);   // low 16 bits = (c2 << 8) | c3
  return ~__byte_perm(lo, hi, 0x5410);                   // a set sign bit = NOT a candidate
}

// When the scene is large and candidates are sparse (swept operand), most 32-sphere words are empty for ALL 32
// rays of a warp (dense 10 k scene: ~80 %): first reduce the 32 accumulators with integer max (full issue
// rate; an accumulator read as a signed integer is >= 0 iff its sign bit is clear) and pack only if some lane
// of the warp has a candidate in this word.  Same bits as sign_word in every case.
__device__ __forceinline__ int max_of(const int (&v)[32]) {
  int m[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    m[k] = v[8 * k];
#pragma unroll
    for (int i = 1; i < 8; ++i) m[k] = max(m[k], v[8 * k + i]);
  }
  return max(max(m[0], m[1]), max(m[2], m[3]));
}
// word of v if `any` (warp-uniform: some lane of the warp has a candidate in it), else 0
__device__ __forceinline__ uint32_t sign_word_if(bool any, const int (&v)[32]) { return any ? sign_word(v) : 0u; }

// Hierarchical form of the same word.  Candidates are sparse AND the 32 rays of a warp are neighbours, so most
// 8-sphere octets are empty for the whole warp (final scene: ~90 %): reduce every octet with full-rate
// integer max (VIMNMX3; an accumulator read as a signed integer is >= 0 iff its sign bit is clear), let the warp
// vote per word and then per octet, and run the half-rate funnel-shift packing only for octets in which some
// lane has a candidate.  Same bits as sign_word in every case; WARP-CONVERGENT.
__device__ __forceinline__ int max_of8(const int* v) {
  return max(max(max(v[0], v[1]), v[2]), max(max(max(v[3], v[4]), v[5]), max(v[6], v[7])));
}
__device__ __forceinline__ uint32_t sign_word_oct(const int (&v)[32]) {
  int o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = max_of8(&v[8 * k]);
  if (!__any_sync(0xffffffffu, max(max(o[0], o[1]), max(o[2], o[3])) >= 0)) return 0u;
  uint32_t c[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t m = 0xffu;                                   // all sign bits set = no candidate
    if (__any_sync(0xffffffffu, o[k] >= 0)) {
      m = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) m = __funnelshift_l((uint32_t)v[8 * k + i], m, 1);
    }
    c[k] = m;
  }
  const uint32_t hi = __byte_perm(c[1], c[0], 0x0040);
  const uint32_t lo = __byte_perm(c[3], c[2], 0x0040);
  return ~__byte_perm(lo, hi, 0x5410);
}
// how the scan warps turn 32 accumulators into a bitmap word: 0 = always pack (sign_word), 1 = skip words that are
// empty for the whole warp, 2 = skip empty words and, inside the others, empty octets
#ifndef UECRT_TC_PACK
#define UECRT_TC_PACK 2
#endif
__device__ __forceinline__ uint32_t pack_word(const int (&v)[32]) {
#if UECRT_TC_PACK == 0
  return sign_word(v);
#elif UECRT_TC_PACK == 1
  return sign_word_if(__any_sync(0xffffffffu, max_of(v) >= 0), v);
#else
  return sign_word_oct(v);
#endif
}

// SCAN: one thread of the scan warpgroup (threadIdx.x in [kShadeThreads, kShadeThreads + 128)).
template <bool SWEEP>
__device__ __forceinline__ void scanner_role(unsigned char* smem_raw, uint32_t tmem_base, int nchunks, uint32_t npasses, int fixed_rounds,
                                             unsigned long long* prof_counters = nullptr) {
#ifdef UECRT_TC_PROFILE
  long long pw = 0, pl = 0, ps = 0, plast = clock64();
  unsigned long long ptiles = 0, p_ready_now = 0, p_blocked = 0;
#define UECRT_TCS(acc) { const long long now_ = clock64(); acc += now_ - plast; plast = now_; }
#else
#define UECRT_TCS(acc)
#endif
  Header* h = reinterpret_cast<Header*>(smem_raw);
  const Layout lay = make_layout((uint32_t)nchunks * kNC, SWEEP ? 2u : 1u);
  const uint32_t st = threadIdx.x - kShadeThreads, sg = st >> 7, row = st & 127u, lane = row & 31u;
  uint32_t lane_taddr = tmem_base + ((row & ~31u) << 16);
  uint32_t full0 = smem_addr(&h->tmem_full[0]), empty0 = smem_addr(&h->tmem_empty[0]), done0 = smem_addr(&h->scan_done[0]);
  // this ray's bitmap word 0 of group 0 / its summary word, as shared addresses (word w of group g: + (w * kShadeThreads + g * 128) * 4)
  uint32_t bitmap0 = smem_addr(smem_raw + lay.bitmap) + row * 4u;
  uint32_t summary0 = smem_addr(smem_raw + lay.summary) + (sg * kShadeThreads + row) * 4u;
  pin(lane_taddr); pin(full0); pin(empty0); pin(done0); pin(bitmap0); pin(summary0);
  uint32_t seq = 0, slot = 0, parity = 0;                     // slot = seq % kRing, parity = (seq / kRing) & 1, kept incrementally
  for (PassIter it; it.begin(h, fixed_rounds); it.next(npasses)) {
    for (uint32_t g = 0; g < (uint32_t)kGroups; ++g) {
      uint32_t o = bitmap0 + g * (kGroupThreads * 4u);        // word w of this ray at o + w * kShadeThreads * 4
      uint32_t sum = 0;                                       // bit (31 - w): word w is non-zero
      for (int j = 0; j < nchunks; ++j, ++seq, o += (kNC / 32) * kShadeThreads * 4u, slot = slot == kRing - 1 ? 0u : slot + 1u, parity ^= (slot == 0u ? 1u : 0u)) {
        if (kScanGroups > 1 && (uint32_t)(j % kScanGroups) != sg) continue;      // another warpgroup's chunk
        const uint32_t use = parity;
#ifdef UECRT_TC_PROFILE
        if (mbar_test(full0 + 8u * slot, use & 1u)) { ++p_ready_now; } else { const long long b0_ = clock64(); mbar_wait(full0 + 8u * slot, use & 1u); ++p_blocked; if (prof_counters && lane == 0 && j < 4) { atomicAdd(prof_counters + 26 + j, 1ull); if (j == 0) { atomicAdd(prof_counters + 42 + g, 1ull); atomicAdd(prof_counters + 46 + g, (unsigned long long)(clock64() - b0_)); } } }
#else
        mbar_wait(full0 + 8u * slot, use & 1u);
#endif
        fence_after_sync();
        UECRT_TCS(pw)
#ifdef UECRT_TC_TRACE
        const bool tr_ = prof_counters && blockIdx.x == 0 && st == 32u && seq >= 3000u && seq < 3064u;
        if (tr_) prof_counters[128 + 8 * (seq - 3000u) + 2] = clock64();
#endif
        const uint32_t t0 = lane_taddr + slot * kNC;
        int va[32], vb[32];
#ifdef UECRT_TC_TRACE
        const bool ph_ = prof_counters && blockIdx.x == 0 && lane == 0 && seq >= 3000u && seq < 3256u;
        long long ph0_ = 0, ph1_ = 0, ph2_ = 0, ph3_ = 0, ph4_ = 0;
        if (ph_) ph0_ = clock64();
#endif
        tmem_ld32(t0, va);
        tmem_ld32(t0 + 32, vb);
        tmem_ld_wait();
#ifdef UECRT_TC_TRACE
        if (ph_) ph1_ = clock64();
#endif
        uint32_t w0, w1, w2, w3;
        w0 = pack_word(va);
        tmem_ld32(t0 + 64, va);
        w1 = pack_word(vb);
        tmem_ld32(t0 + 96, vb);
        sts32(o, w0);
        sts32(o + kShadeThreads * 4u, w1);
#ifdef UECRT_TC_TRACE
        if (ph_) ph2_ = clock64();
#endif
        tmem_ld_wait();
#ifdef UECRT_TC_TRACE
        if (ph_) ph3_ = clock64();
#endif
        fence_before_sync();                              // the whole slot is in registers: hand it back
        if (lane == 0) mbar_arrive(empty0 + 8u * slot);
#ifdef UECRT_TC_TRACE
        if (tr_) prof_counters[128 + 8 * (seq - 3000u) + 3] = clock64();
        if (prof_counters && blockIdx.x == 0 && (st == 0u || st == 96u) && seq >= 3000u && seq < 3064u) prof_counters[128 + 8 * (seq - 3000u) + (st == 0u ? 6 : 7)] = clock64();
#endif
        w2 = pack_word(va);
        w3 = pack_word(vb);
        sts32(o + 2u * kShadeThreads * 4u, w2);
        sts32(o + 3u * kShadeThreads * 4u, w3);
#ifdef UECRT_TC_TRACE
        if (ph_) {   // per scan warp: sums of the phase lengths over 256 chunks
          ph4_ = clock64();
          unsigned long long* p_ = prof_counters + 900 + 8 * (st >> 5);
          atomicAdd(p_ + 0, (unsigned long long)(ph1_ - ph0_)); atomicAdd(p_ + 1, (unsigned long long)(ph2_ - ph1_));
          atomicAdd(p_ + 2, (unsigned long long)(ph3_ - ph2_)); atomicAdd(p_ + 3, (unsigned long long)(ph4_ - ph3_)); atomicAdd(p_ + 4, 1ull);
        }
#endif
        const uint32_t nz = (w0 ? 8u : 0u) | (w1 ? 4u : 0u) | (w2 ? 2u : 0u) | (w3 ? 1u : 0u);
        sum |= nz << (28 - 4 * j);
#ifdef UECRT_TC_TRACE
        if (tr_) prof_counters[128 + 8 * (seq - 3000u) + 4] = clock64();
#endif
        UECRT_TCS(pl)
      }
      sts32(summary0 + g * (kGroupThreads * 4u), sum);
      mbar_arrive(done0 + 8u * g);
#ifdef UECRT_TC_PROFILE
      ++ptiles;
      UECRT_TCS(ps)
#endif
    }
  }
#ifdef UECRT_TC_PROFILE
  if (prof_counters && lane == 0) {
    atomicAdd(prof_counters + 12, ptiles);
    atomicAdd(prof_counters + 13, (unsigned long long)pw);
    atomicAdd(prof_counters + 14, (unsigned long long)pl);
    atomicAdd(prof_counters + 15, (unsigned long long)ps);
    atomicAdd(prof_counters + 32 + ((threadIdx.x >> 5) & 3), (unsigned long long)pw);
    atomicAdd(prof_counters + 36 + ((threadIdx.x >> 5) & 3), (unsigned long long)pl);
    atomicAdd(prof_counters + 24, p_ready_now);
    atomicAdd(prof_counters + 25, p_blocked);
  }
#endif
}

}  // namespace tc
}  // namespace uecrt
