// tc_collect.cuh -- the SCAN role and the shade-side hand-off of the tensor-core scan.
//
// The MMAs leave G = f - P for every (ray, sphere) pair in TMEM (tc_scan.cuh); a sphere may be hit iff the
// sign bit of its word is clear.  Turning 128 x 512 accumulators per tile into something a shade thread can
// walk is what paces the whole kernel: it is integer work on the ALU pipe, where shifts and logic ops issue
// at HALF rate and only integer min/max at full rate (profiles/iops_r1.log).  Round 1 packed one bit per
// sphere (one funnel shift per accumulator: 2.3 k cycles per tile on the one scan warp of a sub-partition,
// twice what the tile's MMAs take).  Now the scan warps reduce every QUAD of spheres with 3-input integer max
// (an accumulator read as a signed integer is >= 0 iff its sign bit is clear, so the max of four is >= 0 iff
// one of them may be hit: 2 full-rate instructions per quad) and pack one bit per quad (one half-rate
// shift): half the pipe cycles and 3/4 of the issue slots per accumulator, still branch-free straight-line code.
// (Skipping words or octets that are empty for all 32 rays of a warp costs a vote + branch each and made the
// single-warp scan stream latency-bound: 19.9 / 26.1 ms against 18.7, profiles/README.md.)
// The shade thread walks the set bits in ascending order; for a flagged quad it re-evaluates the four
// discriminants with the FP32 scan's packed form and bound (closest_hit.cuh: 7 packed FP32 instructions per two
// spheres on the otherwise idle FMA pipe; {f >= P} is again a superset of the reference's "discriminant >= 0")
// and applies the per-candidate test of the resolve phase (test_fast / test_exact: reference root selection,
// closed interval, later index wins -- hittable_list.hpp:32-58, sphere.hpp:32-39) to the survivors, in index order.
#pragma once
#include "closest_hit.cuh"
#include "tc_scan.cuh"

namespace uecrt {
namespace tc {

// SHADE: publish this thread's ray for the current round: its 32 K slots go to TMEM lane t of the group's
// A tile (warp-collective store).  A thread without a live ray passes live = false: its row is still
// written (every row of the tile takes part in the MMA) with P = +1e30, so all its accumulators
// come out negative and its bitmap is empty.
__device__ __forceinline__ void shade_next_pass(const Shade& s);
__device__ __forceinline__ void shade_submit(const Shade& s, const ScanRay& ray, bool live) {
  const float a[kTerms] = {ray.dx * ray.dx, ray.dy * ray.dy, ray.dz * ray.dz, 2.0f * ray.dx * ray.dy, 2.0f * ray.dx * ray.dz,
                           2.0f * ray.dy * ray.dz, ray.nqx, ray.nqy, ray.nqz, 1.0f};
  float row[kSlots];
  fill_row_a(a, live ? ray.p_low : 1e30f, row);
  __syncwarp();
  tmem_st32(s.a_taddr, row);
  tmem_st_wait();
  fence_before_sync();
  shade_next_pass(s);
}
// SHADE: wait for the bitmaps of pass m (tc_scan.cuh PassIter: m = the round when the operand is resident)
__device__ __forceinline__ void shade_wait(const Shade& s, uint32_t m) {
  if (UECRT_TC_NAMED_BARRIERS) nbar_sync(s.scan_done, kBarScanDoneCount);
  else mbar_wait(s.scan_done, m & 1u);
}
// SHADE, swept operand: this thread has consumed its bitmap of the current pass; the ray tile in TMEM stays
// as it is and takes part in the next pass (the arrive releases the bitmap to the scan warps: they only
// write it again behind MMAs the issuer launches after this barrier's phase completes)
__device__ __forceinline__ void shade_next_pass(const Shade& s) {
  if (UECRT_TC_NAMED_BARRIERS) { nbar_arrive(s.a_ready, kBarAReadyCount); return; }
  __syncwarp();
  if ((threadIdx.x & 31u) == 0u) mbar_arrive(s.a_ready);    // one arrival per warp (the barrier counts warps)
}

// SHADE: f(bit index within the pass) for every set bit of this thread's ray, ascending (bit index = sphere index with one
// bit per sphere, quad index with one bit per quad).  WARP-CONVERGENT: all 32 lanes must call it (a lane without a live ray
// passes live = false).  Every trip of the loop hands one bit to every lane that still has one (a lane whose current word is
// exhausted moves on to its next non-empty word: through the summary word with one bit per sphere, word by word otherwise),
// so the warp makes max-over-lanes(set bits) trips and f runs once per trip with the lanes converged.
template <class F>
__device__ __forceinline__ void for_each_flag(Shade& s, uint32_t buf, bool live, F&& f) {
  const uint32_t* bm = s.bitmap + buf * (uint32_t)((kScanGroups + s.nwords) * kShadeThreads);
  uint32_t m = 0u;
  int w = 0;
  if (kQuad == 1) {
    uint32_t sum = 0u;
    if (live) {
#pragma unroll
      for (int sg = 0; sg < kScanGroups; ++sg) sum |= bm[sg * kShadeThreads];
    }
    for (;;) {
      if (m == 0u && sum != 0u) {
        w = __clz((int)sum);
        sum &= ~(0x80000000u >> w);
        m = bm[(kScanGroups + w) * kShadeThreads];
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
  } else {
    w = live ? -1 : s.nwords;
    for (;;) {
      while (m == 0u && w + 1 < s.nwords) { ++w; m = bm[(kScanGroups + w) * kShadeThreads]; }
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
}

// debugging / tests: this ray's bitmap words of the current pass, as the scan warps left them
__device__ __forceinline__ void dump_flags(const Shade& s, uint32_t buf, uint32_t* out) {
  const uint32_t* bm = s.bitmap + buf * (uint32_t)((kScanGroups + s.nwords) * kShadeThreads);
  for (int w = 0; w < s.nwords; ++w) out[w] = bm[(kScanGroups + w) * kShadeThreads];
}

// Per-ray constants of the quad re-test (the FP32 scan's packed operands, closest_hit.cuh scan_pairs)
struct QuadRay {
  unsigned long long dx2, dy2, dz2, qx2, qy2, qz2;
  int thr;
};
__device__ __forceinline__ QuadRay make_quad_ray(const ScanRay& r) {
  QuadRay q;
  q.dx2 = pk2(r.dx, r.dx); q.dy2 = pk2(r.dy, r.dy); q.dz2 = pk2(r.dz, r.dz);
  q.qx2 = pk2(r.nqx, r.nqx); q.qy2 = pk2(r.nqy, r.nqy); q.qz2 = pk2(r.nqz, r.nqz);
  q.thr = r.thr;
  return q;
}
// quads: per quad four float4 {-cx[4]}, {-cy[4]}, {-cz[4]}, {nk[4]} (scene_pack.hpp pack_scene_quads; nk carries the
// FP32 scan's bound).  Returns a 4-bit mask: bit j = sphere 4*quad + j passes f >= P.
__device__ __forceinline__ uint32_t quad_candidates_v(const ulonglong2 X, const ulonglong2 Y, const ulonglong2 Z, const ulonglong2 K, const QuadRay& r) {
  unsigned long long cd0 = mul2(X.x, r.dx2), cd1 = mul2(X.y, r.dx2);
  unsigned long long e0 = fma2(X.x, r.qx2, K.x), e1 = fma2(X.y, r.qx2, K.y);
  cd0 = fma2(Y.x, r.dy2, cd0); cd1 = fma2(Y.y, r.dy2, cd1);
  e0 = fma2(Y.x, r.qy2, e0); e1 = fma2(Y.y, r.qy2, e1);
  cd0 = fma2(Z.x, r.dz2, cd0); cd1 = fma2(Z.y, r.dz2, cd1);
  e0 = fma2(Z.x, r.qz2, e0); e1 = fma2(Z.y, r.qz2, e1);
  const unsigned long long g0 = fma2(cd0, cd0, e0), g1 = fma2(cd1, cd1, e1);
  int f0, f1, f2, f3;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(f0), "=r"(f1) : "l"(g0));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(f2), "=r"(f3) : "l"(g1));
  return (f0 >= r.thr ? 1u : 0u) | (f1 >= r.thr ? 2u : 0u) | (f2 >= r.thr ? 4u : 0u) | (f3 >= r.thr ? 8u : 0u);
}
__device__ __forceinline__ uint32_t quad_candidates(const float4* __restrict__ quads, int quad, const QuadRay& r) {
  const ulonglong2* q = reinterpret_cast<const ulonglong2*>(quads) + 4 * (size_t)quad;
  return quad_candidates_v(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), r);
}
// the same tables in shared memory (resident operand): 16-byte loads from a shared address
__device__ __forceinline__ ulonglong2 lds_u64x2(uint32_t saddr) {
  ulonglong2 v;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t quad_candidates_s(uint32_t s_quads, int quad, const QuadRay& r) {
  const uint32_t q = s_quads + 64u * (uint32_t)quad;
  return quad_candidates_v(lds_u64x2(q), lds_u64x2(q + 16u), lds_u64x2(q + 32u), lds_u64x2(q + 48u), r);
}
// PLANAR copy of the table: four arrays {-cx[4]}, {-cy[4]}, {-cz[4]}, {nk[4]} of 16 B per quad, `plane` bytes apart.  The
// lanes of a warp look up different quads: with 64 B per quad every lane's 16 bytes fall on one of TWO bank groups
// (ncu: 47 % of the kernel's shared-memory wavefronts were bank-conflict replays), with 16 B per quad on one of eight.
__device__ __forceinline__ uint32_t quad_candidates_sp(uint32_t s_quads, uint32_t plane, int quad, const QuadRay& r) {
  const uint32_t q = s_quads + 16u * (uint32_t)quad;
  return quad_candidates_v(lds_u64x2(q), lds_u64x2(q + plane), lds_u64x2(q + 2u * plane), lds_u64x2(q + 3u * plane), r);
}

// one bit per sphere: 32 accumulators -> one bitmap word, bit (31 - i) = (v[i] >= 0).  Four independent funnel-shift
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
// the max of every quad of 32 accumulators (8 values, full-rate VIMNMX3 + VIMNMX each)
__device__ __forceinline__ void quad_max(const int (&v)[32], int (&q)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) q[k] = max(max(max(v[4 * k], v[4 * k + 1]), v[4 * k + 2]), v[4 * k + 3]);
}
// 8 quad maxima -> 8 flags in the low byte: bit (7 - k) SET = every accumulator of quad k is negative (no candidate).
// Two funnel-shift chains (mask = (mask << 1) | sign) of four, merged by one shift-add.
__device__ __forceinline__ uint32_t quad_signs(const int (&q)[8]) {
  uint32_t a = 0, b = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    a = __funnelshift_l((uint32_t)q[k], a, 1);
    b = __funnelshift_l((uint32_t)q[4 + k], b, 1);
  }
  return a * 16u + b;
}

// Position in the issuer's chunk sequence: passes (PassIter), groups, chunks; slot = seq % kRing and the parity of the
// slot's current use, kept incrementally.
struct ChunkPos {
  PassIter it;
  uint32_t g = 0, slot = 0, parity = 0;
  int j = 0;
  // one chunk further; false when the sequence ends (the termination vote is read at a round's first tile)
  __device__ __forceinline__ bool next(const Header* h, int fixed_rounds, uint32_t npasses, int nchunks) {
    slot = slot == kRing - 1 ? 0u : slot + 1u;
    parity ^= (slot == 0u ? 1u : 0u);
    if (++j < nchunks) return true;
    j = 0;
    if (++g < (uint32_t)kGroups) return true;
    g = 0;
    it.next(npasses);
    return it.begin(h, fixed_rounds);
  }
};

// SCAN: one thread of a scan warpgroup (threadIdx.x in [kShadeThreads, kShadeThreads + kScanThreads)).
// A tcgen05.ld round trip is ~150 cycles (tools/microbench/ldtm_probe.cu) and tcgen05.wait::ld waits for ALL of a
// thread's loads, so the loop is software-pipelined over the chunk sequence: the quad maxima of a 32-column load
// free its registers at once, the next load goes into them while the shifts of the previous one run, and the
// first half of the warp's NEXT chunk is requested (after a non-blocking look at its tmem_full barrier) before this
// chunk's second half is packed.  Chunks are consumed in the issuer's order; with kScanGroups warpgroups, warpgroup
// sg takes the chunks whose sequence number is sg modulo kScanGroups, so every SM sub-partition has kScanGroups scan
// warps whose load round trips, barrier checks and dependent chains overlap.
template <bool SWEEP>
__device__ __forceinline__ void scanner_role(unsigned char* smem_raw, uint32_t tmem_base, int nchunks, uint32_t npasses, int fixed_rounds,
                                             unsigned long long* prof_counters = nullptr) {
  static_assert(kScanGroups == 1 || !UECRT_TC_NAMED_BARRIERS, "the named-barrier hand-off counts one scan warpgroup");
#ifdef UECRT_TC_PROFILE
  long long pw = 0, pl = 0, plast = clock64();
  unsigned long long ptiles = 0, p_ready_now = 0, p_blocked = 0;
#define UECRT_TCS(acc) { const long long now_ = clock64(); acc += now_ - plast; plast = now_; }
#else
#define UECRT_TCS(acc)
#endif
  Header* h = reinterpret_cast<Header*>(smem_raw);
  const Layout lay = make_layout((uint32_t)nchunks * kNC, SWEEP ? 2u : 1u);
  const uint32_t st = threadIdx.x - (uint32_t)kScanFirstThread, sg = st >> 7, row = st & 127u;
  uint32_t is_lane0 = (row & 31u) == 0u ? 1u : 0u;
  uint32_t lane_taddr = tmem_base + ((row & ~31u) << 16);
  uint32_t full0 = smem_addr(&h->tmem_full[0]), empty0 = smem_addr(&h->tmem_empty[0]), done0 = smem_addr(&h->scan_done[0]);
  // this ray's bitmap word of chunk 0, group 0, buffer 0, as a shared address (chunk c of group g: + (c * kShadeThreads + g * 128) * 4)
  uint32_t bitmap0 = smem_addr(smem_raw + lay.bitmap) + row * 4u;
  pin(lane_taddr); pin(full0); pin(empty0); pin(done0); pin(bitmap0); pin(is_lane0);
  ChunkPos pos;
  if (!pos.it.begin(h, fixed_rounds)) return;
  for (uint32_t k = 0; k < sg; ++k)
    if (!pos.next(h, fixed_rounds, npasses, nchunks)) return;
#ifdef UECRT_TC_TRACE
  uint32_t seq = sg;
#endif
  int va[32], vb[32];
  uint32_t sum = 0, tile_key = 0xffffffffu;                 // one bit per sphere: bit (31 - w) = word w of the tile is non-zero (this warpgroup's chunks)
  mbar_wait(full0 + 8u * pos.slot, pos.parity);
  fence_after_sync();
  tmem_ld32(lane_taddr + pos.slot * kNC, va);
  tmem_ld32(lane_taddr + pos.slot * kNC + 32, vb);
  UECRT_TCS(pw)
  for (;;) {
#ifdef UECRT_TC_TRACE
    const bool trs_ = prof_counters && blockIdx.x == 0 && is_lane0 && seq >= 3000u && seq < 3048u;
    unsigned long long* tre_ = prof_counters + 128 + 20 * (seq - 3000u) + 4 + 4 * ((st >> 5) & 3u);
    if (trs_) tre_[0] = clock64();
#endif
    const uint32_t t0 = lane_taddr + pos.slot * kNC;
    const bool first_of_tile = tile_key != (pos.it.m * (uint32_t)kGroups + pos.g);   // this warpgroup's first chunk of the tile
    tile_key = pos.it.m * (uint32_t)kGroups + pos.g;
    int qa[8], qb[8];
    uint32_t c0, c1, c2, c3;                              // one bit per quad: 8 flags each; one bit per sphere: a bitmap word each
    tmem_ld_wait();
    if (kQuad == 4) quad_max(va, qa); else c0 = sign_word(va);
    tmem_ld32(t0 + 64, va);
    if (kQuad == 4) quad_max(vb, qb); else c1 = sign_word(vb);
    tmem_ld32(t0 + 96, vb);
    if (kQuad == 4) { c0 = quad_signs(qa); c1 = quad_signs(qb); }
    // this warp's next chunk
    ChunkPos np = pos;
    bool more = true;
#pragma unroll
    for (int k = 0; k < kScanGroups; ++k) more = more && np.next(h, fixed_rounds, npasses, nchunks);
    const bool ahead = more && __all_sync(0xffffffffu, mbar_test(full0 + 8u * np.slot, np.parity));   // its MMAs are complete already
    tmem_ld_wait();
    fence_before_sync();                                  // the whole slot is in registers: hand it back
    if (UECRT_TC_NAMED_BARRIERS) nbar_arrive(kBarEmpty + pos.slot, kBarEmptyCount);
    else if (is_lane0) mbar_arrive(empty0 + 8u * pos.slot);
#ifdef UECRT_TC_TRACE
    if (trs_) tre_[1] = clock64();
#endif
    const uint32_t nt0 = lane_taddr + np.slot * kNC;
    if (kQuad == 4) quad_max(va, qa); else c2 = sign_word(va);
    if (ahead) { fence_after_sync(); tmem_ld32(nt0, va); }
    if (kQuad == 4) quad_max(vb, qb); else c3 = sign_word(vb);
    if (ahead) tmem_ld32(nt0 + 32, vb);
    // this ray's summary word of the tile's bitmap buffer (resident operand: tile m of a group goes to buffer m & 1); word w follows at + (1 + w) * kShadeThreads * 4
    const uint32_t ob = bitmap0 + pos.g * (kGroupThreads * 4u) + (SWEEP ? 0u : (pos.it.m & 1u) * lay.bitmap_buf_bytes);
    if (kQuad == 4) {
      c2 = quad_signs(qa); c3 = quad_signs(qb);
      const uint32_t hi = __byte_perm(c1, c0, 0x0040);    // low 16 bits = (c0 << 8) | c1
      const uint32_t lo = __byte_perm(c3, c2, 0x0040);    // low 16 bits = (c2 << 8) | c3
      sts32(ob + ((uint32_t)kScanGroups + (uint32_t)pos.j) * (kShadeThreads * 4u), ~__byte_perm(lo, hi, 0x5410));   // a set sign bit = NOT a candidate; bit 31 = quad 0 of the chunk
    } else {
      const uint32_t o = ob + ((uint32_t)kScanGroups + 4u * (uint32_t)pos.j) * (kShadeThreads * 4u);
      sts32(o, c0); sts32(o + kShadeThreads * 4u, c1); sts32(o + 2u * kShadeThreads * 4u, c2); sts32(o + 3u * kShadeThreads * 4u, c3);
      const uint32_t nz = (c0 ? 8u : 0u) | (c1 ? 4u : 0u) | (c2 ? 2u : 0u) | (c3 ? 1u : 0u);
      sum = (first_of_tile ? 0u : sum) | (nz << (28 - 4 * pos.j));
      sts32(ob + sg * (kShadeThreads * 4u), sum);
    }
#ifdef UECRT_TC_TRACE
    if (trs_) tre_[2] = clock64();
    seq += kScanGroups;
#endif
    UECRT_TCS(pl)
    if (UECRT_TC_NAMED_BARRIERS) {
      if (pos.j + 1 == nchunks) { __threadfence_block(); nbar_arrive(kBarScanDone + pos.g, kBarScanDoneCount); }   // the tile's last chunk
    } else {
      __syncwarp();                                       // the 32 lanes' words are ordered before the arrival
      if (is_lane0) mbar_arrive(done0 + 8u * pos.g);      // scan_done counts one arrival per scan warp and chunk of the tile
    }
#ifdef UECRT_TC_PROFILE
    if (pos.j + 1 == nchunks) ++ptiles;
#endif
    if (!more) break;
    pos = np;
    if (!ahead) {
#ifdef UECRT_TC_PROFILE
      ++p_blocked;
#endif
      mbar_wait(full0 + 8u * pos.slot, pos.parity);
      fence_after_sync();
      tmem_ld32(nt0, va);
      tmem_ld32(nt0 + 32, vb);
    }
#ifdef UECRT_TC_PROFILE
    else ++p_ready_now;
#endif
    UECRT_TCS(pw)
  }
#ifdef UECRT_TC_PROFILE
  if (prof_counters && is_lane0) {
    atomicAdd(prof_counters + 12, ptiles);
    atomicAdd(prof_counters + 13, (unsigned long long)pw);
    atomicAdd(prof_counters + 14, (unsigned long long)pl);
    atomicAdd(prof_counters + 32 + ((threadIdx.x >> 5) & 3), (unsigned long long)pw);
    atomicAdd(prof_counters + 36 + ((threadIdx.x >> 5) & 3), (unsigned long long)pl);
    atomicAdd(prof_counters + 24, p_ready_now);
    atomicAdd(prof_counters + 25, p_blocked);
  }
#endif
}

}  // namespace tc
}  // namespace uecrt
