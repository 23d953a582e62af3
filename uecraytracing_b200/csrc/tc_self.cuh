// tc_self.cuh -- tensor-core candidate scan WITHOUT a scan role: every shade warp reads the accumulators of
// its own 32 rays ("self-scan" pipeline; resident sphere operand, K1 only).
//
// The warp-specialised pipeline of tc_scan.cuh / tc_collect.cuh funnels all 128 x 512 accumulators of a tile
// through ONE scan warp per SM sub-partition.  That warp is latency-bound (tcgen05.ld round trip ~190 cycles,
// tcgen05.wait::ld waits for all of a thread's loads; 2.3 k cycles per tile at ~0.2 instructions per cycle) and
// closes a loop with the four shade groups that wait for it: the kernel ran at 0.45 useful instructions per
// cycle and sub-partition with 37 % of all executed instructions being barrier polls (profiles/README.md,
// round 2).  Here the same integer work (3-input max over every quad of accumulators, one funnel shift per
// quad; tc_collect.cuh quad_max / quad_signs) is done by the warp that owns the rays, so
//   * every SM sub-partition runs kGroups identical warps whose TMEM round trips, shared-memory loads and
//     dependent chains overlap each other (thread-level parallelism instead of one pacing warp),
//   * the quad flags never leave the thread: no bitmap hand-off, no scan_done barrier,
//   * the scan warpgroup's registers go to more shade groups (6 x 128 paths at 80 registers).
// Final scene, 1200 x 675 x 500 spp on one B200: 83.5 ms per frame against 88.7 with the scan-warp pipeline of round 2's
// start and 91.9 in round 1; what bounds it now is the length of a group's serial round, see DESIGN.md.
// Roles of a CTA (kThreads = kGroups * 128 + 128, one CTA per SM):
//   SHADE   kGroups groups x 128 threads, one path per thread.  Per round: lifted ray -> row of an A tile in
//           TMEM, then for every 128-sphere chunk wait for its MMAs, read the 128 accumulators of the own ray
//           (4 x tcgen05.ld.32x32b.x32), reduce to 32 quad flags, release the accumulator slot; then walk the
//           flags, resolve the closest hit (same tests as every other scan) and shade.
//   ISSUER  kLanes = 3 warps, one per LANE.  Lane i = accumulator slot i (128 TMEM columns) + A slot i (32 columns);
//           tile T = round * kGroups + group flows through lane T % 3, so with kGroups = 6 lane i serves groups i and
//           i + 3 in turn.  Per tile the lane's issuer waits for the A rows, then per chunk: wait until the group has
//           read the slot, issue 4 tcgen05.mma (K = 32 slots), commit.  A lane is strictly sequential (MMA -> read ->
//           MMA ...); three lanes keep three groups reading while the tensor pipe serves whichever lane is ready.
//           (One issuer over one shared ring, whole tiles in fixed order, left all three slots with the SAME group:
//           one chunk per read latency, 2.98 k cycles per tile.)
// TMEM (512 columns): 3 accumulator slots x 128 columns, 3 A slots x 32 columns.
// Barriers (mbarriers; a parity wait only tells adjacent phases apart, so every barrier's phases are visited in
// order by ONE waiting party, which never starts waiting for phase u before phase u - 1 has completed):
//   a_ready[i]     4 arrivals (warps of the tile's group)       waited by issuer i
//   a_free[i]      tcgen05.commit behind the tile's last MMA    waited by the group of the lane's NEXT tile, T + 3; with
//                  kGroups = 6 that group's previous tile was T - 3 of the same lane, whose a_free phase completed
//                  (thousands of cycles) before it comes here
//   grp_full[g]    tcgen05.commit behind chunk c of group g (c counts the group's chunks, parity c & 1): waited by the
//                  group's four warps; chunk c + 1 is only issued after all four have read chunk c
//   tmem_empty[i]  4 arrivals (warps of the group that read the slot)   waited by issuer i
// Termination: the vote of tc_scan.cuh (done flags per round and warp, read two rounds later), per lane (see lane_done_at).
#pragma once
#include "tc_collect.cuh"

namespace uecrt {
namespace ts {

#ifndef UECRT_TS_GROUPS
#define UECRT_TS_GROUPS 6
#endif
constexpr int kGroups = UECRT_TS_GROUPS;
constexpr int kShadeThreads = kGroups * 128;
constexpr int kThreads = kShadeThreads + 128;          // + the issuer warpgroup (one working warp)
constexpr int kNC = tc::kNC;                           // 128 spheres per chunk
constexpr int kLanes = 3;                              // accumulator slot + A slot + issuer warp each
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kATmemCol = kLanes * kNC;           // A slot i: TMEM columns [384 + 32 i, 384 + 32 i + 32)
static_assert(kATmemCol + kLanes * tc::kSlots <= kTmemCols, "TMEM budget");
static_assert(kGroups == 3 || kGroups == 6, "a lane alternates between the groups g and g + 3 (a_free phase order)");
// launch allocation, redistributed by setmaxnreg: the issuer warpgroup keeps 24 registers, the shade warps take the rest
constexpr int kLaunchRegs = (65536 / kThreads) / 8 * 8;
constexpr int kIssuerRegs = 24;
constexpr int kShadeRegsFit = ((kThreads * kLaunchRegs - 128 * kIssuerRegs) / kShadeThreads) / 8 * 8;
#ifndef UECRT_TS_SHADE_REGS
#define UECRT_TS_SHADE_REGS (uecrt::ts::kShadeRegsFit > 128 ? 128 : uecrt::ts::kShadeRegsFit)
#endif
constexpr int kShadeRegs = UECRT_TS_SHADE_REGS;
constexpr int kMaxRows = 1152;                         // 9 chunks: operand + private flag words + shade tables in 227 KB
constexpr uint32_t kHeaderBytes = 1024;

struct Header {
  uint64_t a_ready[kLanes];
  uint64_t a_free[kLanes];
  uint64_t grp_full[kGroups];
  uint64_t tmem_empty[kLanes];
  uint64_t b_loaded;
  uint32_t tmem_base;
  uint32_t pad_;
  uint8_t done_flags[8][32];   // [round & 7][shade warp]
};
static_assert(sizeof(Header) <= kHeaderBytes, "header");

// smem: [Header][B: rows x 128 B][flag words: rows / 128 per thread, word w of thread t at (w * kShadeThreads + t) * 4]
//       [quad table rows x 16 B][spheres rows x 16 B]
struct Layout {
  uint32_t b, words, quads, geom, total;
};
__host__ __device__ inline Layout make_layout(uint32_t rows) {
  Layout l;
  l.b = kHeaderBytes;
  l.words = l.b + rows * 128u;
  l.quads = l.words + (rows / (uint32_t)kNC) * (uint32_t)tc::kWordsPerChunk * (uint32_t)kShadeThreads * 4u;
  l.geom = l.quads + rows * 16u;
  l.total = l.geom + rows * 16u;
  return l;
}

#ifdef __CUDACC__
__device__ __forceinline__ bool is_issuer_thread() { return (int)threadIdx.x >= kShadeThreads; }

// Tuning knobs (each measured in a same-call A/B on the final scene, 100 spp; DESIGN.md section 4 has the numbers).
#ifndef UECRT_TS_LD64
#define UECRT_TS_LD64 0            // 1: two tcgen05.ld of 32 columns in flight (no difference: the reads are issue-bound, not round-trip-bound)
#endif
#ifndef UECRT_TS_IDLE_SKIP
#define UECRT_TS_IDLE_SKIP 0       // 1: tiles / warps without a live ray skip the MMAs / the reads: 25 us off the drain of a launch, 1.2 % onto every round
#endif
// How each hand-off is waited for.  0: tc::mbar_wait (try_wait + NANOSLEEP.SYNCS: the warp re-polls with 5 instructions on
// every wake-up; a quarter of all executed instructions are such polls); -1: spin on test_wait (no hardware suspend);
// N > 0: test_wait with a timer sleep of N ns in between.  The issuer spins on "slot read" (16.8 against 17.3 ms with the
// suspending wait, 17.0 - 17.1 with 20 - 100 ns sleeps); every other variant of every other wait was slower.
#ifndef UECRT_TS_SLEEP_EMPTY
#define UECRT_TS_SLEEP_EMPTY -1
#endif
#ifndef UECRT_TS_SLEEP_READY
#define UECRT_TS_SLEEP_READY 0
#endif
#ifndef UECRT_TS_SLEEP_AFREE
#define UECRT_TS_SLEEP_AFREE 0
#endif
#ifndef UECRT_TS_SLEEP_FULL
#define UECRT_TS_SLEEP_FULL 0
#endif
template <int NS>
__device__ __forceinline__ void wait_backoff(uint32_t bar, uint32_t parity) {
  if (NS == 0) { tc::mbar_wait(bar, parity); return; }
  if (NS < 0) { while (!tc::mbar_test(bar, parity)) {} return; }      // spin without the hardware suspend
  while (!tc::mbar_test(bar, parity)) __nanosleep(NS);
}

__device__ __forceinline__ uint32_t cta_setup(unsigned char* smem_raw, const void* b_global, uint32_t b_bytes) {
  Header* h = reinterpret_cast<Header*>(smem_raw);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kLanes; ++i) {
      tc::mbar_init(tc::smem_addr(&h->a_ready[i]), 4);
      tc::mbar_init(tc::smem_addr(&h->a_free[i]), 1);
      tc::mbar_init(tc::smem_addr(&h->tmem_empty[i]), 4);
    }
    for (int g = 0; g < kGroups; ++g) tc::mbar_init(tc::smem_addr(&h->grp_full[g]), 1);
    tc::mbar_init(tc::smem_addr(&h->b_loaded), 1);
    for (int i = 0; i < 8 * 32; ++i) (&h->done_flags[0][0])[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((threadIdx.x >> 5) == 0) tc::tmem_alloc(tc::smem_addr(&h->tmem_base), kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (threadIdx.x == 0) tc::bulk_load(tc::smem_addr(smem_raw + kHeaderBytes), b_global, b_bytes, tc::smem_addr(&h->b_loaded));
  return h->tmem_base;
}
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base) {
  tc::fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(tmem_base, kTmemCols);
}

// Termination is per LANE: the groups of a lane (g = lane, lane + 3, ...) take strict turns on it, so they are never more
// than one round apart, and the flags of round r - 2 of all of them are ordered before anything the lane's issuer or
// one of its shade warps does in round r (a_ready arrive -> issuer -> commit -> grp_full).  Lanes are independent of
// each other and end when their own groups have drained the shared work queue.
__device__ __forceinline__ bool lane_done_at(const Header* h, uint32_t lane, uint32_t r) {
  const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(&h->done_flags[r & 7u][0]);
  uint32_t all = 0x01010101u;
#pragma unroll
  for (int g = 0; g < kGroups; g += kLanes) all &= f[lane + g];     // four warp flags per word, one word per group
  return all == 0x01010101u;
}
__device__ __forceinline__ bool round_continues(const Header* h, uint32_t lane, uint32_t r) { return !(r >= 2u && lane_done_at(h, lane, r - 2u)); }
// the same through a SHARED address kept in a register (flags0 = shared address of done_flags[0][0]): the generic-pointer
// form makes ptxas re-derive the shared window base from S2UR SR_CgaCtaId in every round
__device__ __forceinline__ bool round_continues_s(uint32_t flags0, uint32_t lane, uint32_t r) {
  if (r < 2u) return true;
  const uint32_t row = flags0 + 32u * ((r - 2u) & 7u) + 4u * lane;
  uint32_t all = 0x01010101u;
#pragma unroll
  for (int g = 0; g < kGroups; g += kLanes) {
    uint32_t f;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(f) : "r"(row + 4u * (uint32_t)g));
    all &= f;
  }
  return all != 0x01010101u;
}
__device__ __forceinline__ void vote_s(uint32_t my_flag0, uint32_t r, bool warp_done) {   // my_flag0 = flags0 + warp index; one lane calls it
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(my_flag0 + 32u * (r & 7u)), "r"(warp_done ? 1u : 0u) : "memory");
}

// ISSUER of lane `lane`: the whole warp runs the loop and waits convergently, one elected lane issues (tc_scan.cuh issuer_role)
__device__ __forceinline__ void issuer_role(unsigned char* smem_raw, uint32_t tmem_base, int nchunks, uint32_t lane, unsigned long long* prof_counters = nullptr) {
#ifdef UECRT_TC_PROFILE
  long long p_ready = 0, p_empty = 0, p_issue = 0, p_last = clock64();
  unsigned long long p_tiles = 0;
#define UECRT_TSI(acc) { const long long now_ = clock64(); acc += now_ - p_last; p_last = now_; }
#else
#define UECRT_TSI(acc)
#endif
  Header* h = reinterpret_cast<Header*>(smem_raw);
  const uint32_t b_smem = tc::smem_addr(smem_raw + kHeaderBytes);
  uint32_t ready = tc::smem_addr(&h->a_ready[lane]), afree = tc::smem_addr(&h->a_free[lane]), full0 = tc::smem_addr(&h->grp_full[0]),
           empty = tc::smem_addr(&h->tmem_empty[lane]);
  uint32_t d_tmem = tmem_base + lane * (uint32_t)kNC, a_tmem = tmem_base + kATmemCol + lane * (uint32_t)tc::kSlots;
  tc::pin(ready); tc::pin(afree); tc::pin(full0); tc::pin(empty); tc::pin(d_tmem); tc::pin(a_tmem);
  tc::mbar_wait(tc::smem_addr(&h->b_loaded), 0);
  uint32_t aparity = 0;                           // parity of the lane's tile counter
  uint32_t eparity = 1;                           // parity of the lane's PREVIOUS chunk (its read releases the slot)
  bool first = true;                              // the lane's first chunk does not wait
  for (uint32_t T = lane;; T += (uint32_t)kLanes) {
    const uint32_t r = T / (uint32_t)kGroups, g = T - r * (uint32_t)kGroups;
    if (!round_continues(h, lane, r)) break;
    UECRT_TSI(p_issue)
    wait_backoff<UECRT_TS_SLEEP_READY>(ready, aparity);
    UECRT_TSI(p_ready)
    aparity ^= 1u;
    // a tile whose four warps all voted "done" carries no live ray (the votes of round r precede its a_ready arrivals):
    // its chunks are committed without MMAs and its warps skip the reads, so a drained group costs its lane partner --
    // the one still finishing long paths at the end of a launch -- four hand-offs instead of a whole tile
    const bool idle_tile = UECRT_TS_IDLE_SKIP && reinterpret_cast<const volatile uint32_t*>(&h->done_flags[r & 7u][0])[g] == 0x01010101u;
    for (int j = 0; j < nchunks; ++j) {
      UECRT_TSI(p_issue)
      if (!first) wait_backoff<UECRT_TS_SLEEP_EMPTY>(empty, eparity);
      UECRT_TSI(p_empty)
      first = false;
      eparity ^= 1u;
      tc::fence_after_sync();
      if (tc::elect_one()) {
        if (!idle_tile) tc::issue_tile_ts(d_tmem, a_tmem, b_smem + (uint32_t)j * (kNC * 128u), kNC);
        if (j + 1 == nchunks) tc::mma_commit(afree);
        tc::mma_commit(full0 + 8u * g);
      }
      __syncwarp();
    }
#ifdef UECRT_TC_PROFILE
    ++p_tiles;
#endif
  }
#ifdef UECRT_TC_PROFILE
  if (prof_counters && (threadIdx.x & 31u) == 0u) {
    atomicAdd(prof_counters + 20, p_tiles);
    atomicAdd(prof_counters + 21, (unsigned long long)p_ready);
    atomicAdd(prof_counters + 22, (unsigned long long)p_empty);
    atomicAdd(prof_counters + 23, (unsigned long long)p_issue);
  }
#endif
}

// SHADE: per-thread constants of the pipeline
struct Shade {
  uint32_t a_ready0, a_free0, full, empty0;      // shared addresses: a_ready[0], a_free[0], grp_full[g], tmem_empty[0]
  uint32_t lane_taddr;                           // TMEM address of this warp's lane quarter, column 0
  uint32_t words;                                // shared address of this thread's flag word 0 (word w: + w * kShadeThreads * 4)
  uint32_t g;
  int nchunks;
  uint32_t gparity = 0;                          // parity of the group's next chunk
  uint32_t is_lane0;                             // 1 in lane 0 of the warp (kept in a register: SR_TID is ~25 cycles away)
};
__device__ __forceinline__ Shade make_shade(unsigned char* smem_raw, uint32_t tmem_base, int nchunks) {
  Header* h = reinterpret_cast<Header*>(smem_raw);
  const Layout lay = make_layout((uint32_t)nchunks * kNC);
  Shade s;
  s.g = threadIdx.x >> 7;
  s.a_ready0 = tc::smem_addr(&h->a_ready[0]); s.a_free0 = tc::smem_addr(&h->a_free[0]);
  s.full = tc::smem_addr(&h->grp_full[s.g]); s.empty0 = tc::smem_addr(&h->tmem_empty[0]);
  s.lane_taddr = tmem_base + ((((threadIdx.x >> 5) & 3u) * 32u) << 16);
  s.words = tc::smem_addr(smem_raw + lay.words) + threadIdx.x * 4u;
  s.nchunks = nchunks;
  s.is_lane0 = (threadIdx.x & 31u) == 0u ? 1u : 0u;
  return s;
}

// SHADE: publish this thread's ray as a row of tile T's A slot (tc_collect.cuh shade_submit)
__device__ __forceinline__ void submit(const Shade& s, uint32_t T, const ScanRay& ray, bool live) {
  const float a[tc::kTerms] = {ray.dx * ray.dx, ray.dy * ray.dy, ray.dz * ray.dz, 2.0f * ray.dx * ray.dy, 2.0f * ray.dx * ray.dz,
                               2.0f * ray.dy * ray.dz, ray.nqx, ray.nqy, ray.nqz, 1.0f};
  float row[tc::kSlots];
  tc::fill_row_a(a, live ? ray.p_low : 1e30f, row);
  const uint32_t aslot = T % (uint32_t)kLanes, ause = T / (uint32_t)kLanes;
  if (T >= (uint32_t)kLanes) {                    // the MMAs of the lane's previous tile have completed
    wait_backoff<UECRT_TS_SLEEP_AFREE>(s.a_free0 + 8u * aslot, (ause - 1u) & 1u);
    tc::fence_after_sync();
  }
  __syncwarp();
  tc::tmem_st32(s.lane_taddr + kATmemCol + aslot * (uint32_t)tc::kSlots, row);
  tc::tmem_st_wait();
  tc::fence_before_sync();
  __syncwarp();
  if (s.is_lane0) tc::mbar_arrive(s.a_ready0 + 8u * aslot);
}

// SHADE: read the accumulators of this thread's ray for every chunk of tile T as the MMAs complete and leave one
// flag word per chunk (bit 31 - k = some sphere of quad k of the chunk may be hit) in the thread's private words.
__device__ __forceinline__ void scan_tile(Shade& s, uint32_t T, bool warp_idle, long long* prof_wait = nullptr) {
  const uint32_t slot = T % (uint32_t)kLanes;
  const bool lane0 = s.is_lane0 != 0u;
  for (int j = 0; j < s.nchunks; ++j) {
#ifdef UECRT_TC_PROFILE
    const long long pw0 = clock64();
#endif
    wait_backoff<UECRT_TS_SLEEP_FULL>(s.full, s.gparity);
#ifdef UECRT_TC_PROFILE
    if (prof_wait) *prof_wait += clock64() - pw0;
#endif
    s.gparity ^= 1u;
    tc::fence_after_sync();
    if (UECRT_TS_IDLE_SKIP && warp_idle) {                              // no live ray in this warp (warp-uniform): nothing to read, hand the slot back
      tc::fence_before_sync();
      __syncwarp();
      if (lane0) tc::mbar_arrive(s.empty0 + 8u * slot);
      continue;
    }
    const uint32_t t0 = s.lane_taddr + slot * (uint32_t)kNC;
    uint32_t c0, c1, c2, c3;
#if UECRT_TS_LD64
    // two 32-column loads in flight: two TMEM round trips per chunk instead of four (64 registers of accumulators)
    int v[32], w[32], q[8], p[8];
    tc::tmem_ld32(t0, v);
    tc::tmem_ld32(t0 + 32, w);
    tc::tmem_ld_wait();
    tc::quad_max(v, q);
    tc::quad_max(w, p);
    tc::tmem_ld32(t0 + 64, v);
    tc::tmem_ld32(t0 + 96, w);
    c0 = tc::quad_signs(q);
    c1 = tc::quad_signs(p);
    tc::tmem_ld_wait();
    tc::fence_before_sync();                      // the whole slot is in registers: hand it back
    __syncwarp();
    if (lane0) tc::mbar_arrive(s.empty0 + 8u * slot);
    tc::quad_max(v, q);
    tc::quad_max(w, p);
    c2 = tc::quad_signs(q);
    c3 = tc::quad_signs(p);
#else
    int v[32], q[8];
    if (tc::kQuad == 1) {                          // one flag per SPHERE (-DUECRT_TC_QUAD=1): four words per chunk, no quad re-test in the resolve
      tc::tmem_ld32(t0, v);
      tc::tmem_ld_wait();
      c0 = tc::sign_word(v);
      tc::tmem_ld32(t0 + 32, v);
      tc::tmem_ld_wait();
      c1 = tc::sign_word(v);
      tc::tmem_ld32(t0 + 64, v);
      tc::tmem_ld_wait();
      c2 = tc::sign_word(v);
      tc::tmem_ld32(t0 + 96, v);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane0) tc::mbar_arrive(s.empty0 + 8u * slot);
      c3 = tc::sign_word(v);
      const uint32_t o = s.words + 4u * (uint32_t)j * (kShadeThreads * 4u);
      tc::sts32(o, c0); tc::sts32(o + kShadeThreads * 4u, c1); tc::sts32(o + 2u * kShadeThreads * 4u, c2); tc::sts32(o + 3u * kShadeThreads * 4u, c3);
      continue;
    }
    tc::tmem_ld32(t0, v);
    tc::tmem_ld_wait();
    tc::quad_max(v, q);
    tc::tmem_ld32(t0 + 32, v);
    c0 = tc::quad_signs(q);
    tc::tmem_ld_wait();
    tc::quad_max(v, q);
    tc::tmem_ld32(t0 + 64, v);
    c1 = tc::quad_signs(q);
    tc::tmem_ld_wait();
    tc::quad_max(v, q);
    tc::tmem_ld32(t0 + 96, v);
    c2 = tc::quad_signs(q);
    tc::tmem_ld_wait();
    tc::fence_before_sync();                      // the whole slot is in registers: hand it back
    __syncwarp();
    if (lane0) tc::mbar_arrive(s.empty0 + 8u * slot);
    tc::quad_max(v, q);
    c3 = tc::quad_signs(q);
#endif
    const uint32_t hi = __byte_perm(c1, c0, 0x0040);    // low 16 bits = (c0 << 8) | c1
    const uint32_t lo = __byte_perm(c3, c2, 0x0040);    // low 16 bits = (c2 << 8) | c3
    tc::sts32(s.words + (uint32_t)j * (kShadeThreads * 4u), ~__byte_perm(lo, hi, 0x5410));   // a set sign bit = NOT a candidate
  }
}

// SHADE: f(quad index) for every set flag of this thread's ray, ascending.  WARP-CONVERGENT (tc_collect.cuh for_each_flag):
// the warp makes max-over-lanes(set flags) trips and f runs once per trip with the lanes converged.
#ifndef UECRT_TS_UNIFORM_TRIPS
#define UECRT_TS_UNIFORM_TRIPS 1
#endif
template <class F>
__device__ __forceinline__ void for_each_flag(const Shade& s, bool live, F&& f) {
  const int nwords = s.nchunks * tc::kWordsPerChunk;
  if (UECRT_TS_UNIFORM_TRIPS && nwords <= 4) {
    // up to four flag words: all loaded at once (independent loads), the trip count is the warp's maximum of the
    // lanes' flag counts (one REDUX), and every trip pops the lane's next flag from registers
    uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
    if (live) {
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(s.words));
      if (nwords > 1) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(s.words + 1u * (kShadeThreads * 4u)));
      if (nwords > 2) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w2) : "r"(s.words + 2u * (kShadeThreads * 4u)));
      if (nwords > 3) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w3) : "r"(s.words + 3u * (kShadeThreads * 4u)));
    }
    const int trips = __reduce_max_sync(0xffffffffu, __popc(w0) + __popc(w1) + __popc(w2) + __popc(w3));
    for (int t = 0; t < trips; ++t) {
      // the lane's lowest flag index: first non-empty word, highest set bit
      const uint32_t m = w0 ? w0 : w1 ? w1 : w2 ? w2 : w3;
      const int base = w0 ? 0 : w1 ? 32 : w2 ? 64 : 96;
      if (m != 0u) {
        const int b = __clz((int)m);
        const uint32_t cleared = m & ~(0x80000000u >> b);
        if (w0) w0 = cleared; else if (w1) w1 = cleared; else if (w2) w2 = cleared; else w3 = cleared;
        f(base + b);
      }
    }
    return;
  }
  uint32_t m = 0u;
  int w = live ? -1 : nwords;
  for (;;) {
    while (m == 0u && w + 1 < nwords) {
      ++w;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m) : "r"(s.words + (uint32_t)w * (kShadeThreads * 4u)));
    }
    if (!__any_sync(0xffffffffu, m != 0u)) return;
    if (m != 0u) {
      const int b = __clz((int)m);
      m &= ~(0x80000000u >> b);
      f(w * 32 + b);
    }
  }
}
#endif  // __CUDACC__

}  // namespace ts
}  // namespace uecrt
