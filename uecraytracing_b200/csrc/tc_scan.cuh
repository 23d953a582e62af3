// tc_scan.cuh -- tensor-core (tcgen05 / TMEM) form of the closest-hit candidate scan.
//
// The scan of closest_hit.cuh evaluates, for every (ray, sphere) pair,
//     f = (c.d)^2 + q.c - k          (candidate iff f >= P; sphere.hpp:27-32 with |d| = 1)
// which is BILINEAR in a lifted ray vector and a lifted sphere vector:
//     f = sum_{i<=j} (d_i d_j)(c_i c_j) + q.c + 1*(-k)
//       = a . b,   a = [dx^2, dy^2, dz^2, 2dxdy, 2dxdz, 2dydz, -qx, -qy, -qz, 1]      (per ray)
//                  b = [cx^2, cy^2, cz^2, cxcy,  cxcz,  cycz,  -cx, -cy, -cz, -k]     (per sphere)
// so the all-pairs test of 128 rays against N spheres is a 128 x N x 10 matrix product.  It runs on
// the 5th-generation tensor cores as tcgen05.mma.kind::tf32 with FP32 accumulation in TMEM.  TF32
// keeps 11 significand bits, so every operand is split a = a_hi + a_lo (each exactly representable
// in TF32) and the product is taken as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi: 3 x 10 = 30 of the 32
// K slots of four K=8 MMA steps; the last two slots carry -P (hi, lo) against a sphere-side 1, so
// the accumulator holds G = f - P and a sphere is a candidate iff G >= 0, i.e. iff the SIGN BIT of
// the TMEM word is clear.  All products of TF32 values are exact in FP32; what is lost is
// a_lo*b_lo (<= 2^-22 |a||b|), the rounding of the lo parts (<= 2^-22 each) and the accumulator's
// own rounding / truncation (bounded in scene_pack.hpp).  The bound on the sum of those is applied exactly like the FP32 scan's bound
// (scene_pack.hpp): the sphere part raises -k, the ray part lowers P, so {f >= P} stays a SUPERSET
// of the reference's "discriminant >= 0" and the resolve phase (closest_hit.cuh) is unchanged.
//
// The RAY operand (A) lives in TMEM: thread t of a shade group writes its 32 slots to TMEM lane t with
// one tcgen05.st.32x32b.x32 (no shared-memory staging, no proxy fence), and the MMAs read it from
// there, so the only shared-memory traffic of an MMA is its 4 KB slice of the sphere operand
// (with both operands in shared memory an M=128, N=128, K=8 TF32 MMA reads 8 KB = 64 cycles of the
// SM's 128 B/clk, i.e. the tensor pipe was shared-memory-bound: 107 cycles per MMA against 74 with A in TMEM,
// tools/microbench/tc_probe.cu T4, profiles/tc_probe_mma_patterns_r1.log).
// Shared-memory layout of the SPHERE operand (K-major, no swizzle; the UMMA "interleaved" canonical layout):
//   a row (sphere) has 32 FP32 slots = 8 chunks of 16 B; rows are grouped by 8;
//   byte offset of (row, chunk) = (row / 8) * 1024 + chunk * 128 + (row % 8) * 16
//   -> core matrix = 8 rows x 16 B contiguous (128 B); LBO (next core matrix along K) = 128 B;
//      SBO (next 8-row group) = 1024 B.  One MMA step consumes 8 slots = 2 chunks.
// TMEM: D tile = 128 lanes (rays) x kNC columns (spheres) of FP32; thread t of a warp w with w % 4 == q
// reads lane 32q + t with tcgen05.ld.32x32b: one thread reads the accumulators of one ray.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#else                       // host-only users (tests/cpp/test_tc_pack.cpp): the packing helpers need no CUDA header
#define __host__
#define __device__
#define __forceinline__ inline
struct { unsigned x; } static threadIdx;
#endif
#include <stdint.h>

namespace uecrt {
namespace tc {

constexpr int kSlots = 32;                        // K slots per row (30 used)
constexpr int kTerms = 10;                        // lifted dimension
constexpr uint32_t kRowGroupBytes = 8u * kSlots * 4u;   // 1024
constexpr uint32_t kLBO = 128u;
constexpr uint32_t kSBO = kRowGroupBytes;

__host__ __device__ inline uint32_t row_chunk_offset(uint32_t row, uint32_t chunk) {
  return (row >> 3) * kRowGroupBytes + chunk * 128u + (row & 7u) * 16u;
}

// hi = v rounded to nearest TF32 (11 significant bits), lo = (v - hi) rounded to nearest TF32 as well.
// v - hi is exact in FP32 (hi is v with the low bits rounded away) and |v - hi| <= 2^-11 |v|; rounding it to 11 bits
// leaves |v - hi - lo| <= 2^-11 * 2^-11 |v| = 4u |v| (u = 2^-24; truncating it, as round 1 did, leaves 8u |v|).
__host__ __device__ inline void split_tf32(float v, float& hi, float& lo) {
#ifdef __CUDA_ARCH__
  const uint32_t b = __float_as_uint(v);
  hi = __uint_as_float((b + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float((__float_as_uint(v - hi) + 0x1000u) & 0xffffe000u);
#else
  union { float f; uint32_t u; } x, y;
  x.f = v; x.u = (x.u + 0x1000u) & 0xffffe000u; hi = x.f;
  y.f = v - hi; y.u = (y.u + 0x1000u) & 0xffffe000u; lo = y.f;
#endif
}

// K-slot order of one row.  A (ray) side: [hi(10) | hi(10) | lo(10) | -P_hi -P_lo];  B (sphere) side:
// [hi(10) | lo(10) | hi(10) | 1 1]  ->  sum over slots = hi.hi + hi.lo + lo.hi - P.
__host__ __device__ inline void fill_row_a(const float (&a)[kTerms], float p, float (&row)[kSlots]) {
  for (int i = 0; i < kTerms; ++i) {
    float h, l;
    split_tf32(a[i], h, l);
    row[i] = h; row[kTerms + i] = h; row[2 * kTerms + i] = l;
  }
  float ph, pl;
  split_tf32(p, ph, pl);
  row[30] = -ph; row[31] = -pl;
}
__host__ __device__ inline void fill_row_b(const float (&b)[kTerms], float (&row)[kSlots]) {
  for (int i = 0; i < kTerms; ++i) {
    float h, l;
    split_tf32(b[i], h, l);
    row[i] = h; row[kTerms + i] = l; row[2 * kTerms + i] = h;
  }
  row[30] = 1.f; row[31] = 1.f;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory matrix descriptor (sm_100): start address, LBO, SBO in 16-byte units,
// version = 1, no swizzle, base offset 0.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
}
// UMMA instruction descriptor: D = F32, A = B = TF32, both K-major, dense, shape M x N.
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// both operands in shared memory (the probe's layout test; the product keeps A in TMEM, see below)
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in TMEM (lane = row, one 32-bit column per K element)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all tcgen05 operations issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// true in exactly one lane of a CONVERGED warp.  ptxas knows the code under it runs in a single thread,
// so tcgen05.mma operands go to uniform registers directly; under a plain `if (lane == 0)` it wraps
// every MMA in a divergence loop (R2UR.BROADCAST + BRA.U.ANY), measured at 208 cycles per MMA against
// 74 with elect.sync (M=128, N=128, K=8 TF32, A in TMEM; tools/microbench/tc_probe.cu T4).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one full warp
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Blocking wait for the phase with the given parity.  The suspend-time hint lets the hardware keep the
// warp asleep until the phase completes: with the default (short) time limit the waiting warps of this
// pipeline executed 43% of all issued instructions just re-polling (ncu source view, profiles/README.md).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(bar), "r"(parity), "r"(1000000u)
      : "memory");
}
// Hardware named barriers for the warp-to-warp hand-offs (a_ready, scan_done, tmem_empty): a warp blocked in bar.sync issues
// nothing, whereas a warp in an mbarrier try_wait loop re-polls (SYNCS + NANOSLEEP.SYNCS + BRA): in round 1 and early round 2
// that polling was 38 % of all executed instructions, 40 % of the issue slots, and NANOSLEEP runs on the quarter-rate XU
// pipe that the shade warps need for MUFU / FLO / conversions (ncu: xu pipe oversubscribed).  Producers bar.arrive, the
// consumer bar.sync; `count` = producer threads + consumer threads.  Every generation of a barrier completes before the next
// one's first arrival (the pipeline's data dependencies guarantee it, see the role functions).  mbarriers remain where the
// async proxy signals completion (tcgen05.commit -> tmem_full, cp.async.bulk -> b_loaded / b_full).
__device__ __forceinline__ void nbar_sync(uint32_t id, uint32_t count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(uint32_t id, uint32_t count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {      // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Keep a loop-invariant value in its register: without this ptxas re-derives shared addresses inside the
// pipeline loops from S2R (SR_CgaCtaId, SR_TID) -- two ~25-cycle special-register reads and a dependent
// integer chain in front of every chunk's barrier wait and first TMEM load (timeline trace: ~130 cycles per chunk).
__device__ __forceinline__ void pin(uint32_t& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// named barrier of `nthreads` threads (tools/microbench/tc_probe.cu only)
__device__ __forceinline__ void group_bar(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 32 consecutive TMEM columns of this thread's lane -> 32 registers (no wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 registers -> 32 consecutive TMEM columns of this thread's lane (warp-collective)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
      "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
      "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
      "f"(v[30]), "f"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// write one 32-slot row (8 x 16 B) into the canonical layout; `base` = shared address of the
// operand tile, row = row within the tile
__device__ __forceinline__ void store_row(unsigned char* base, uint32_t row, const float (&r)[kSlots]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    *reinterpret_cast<float4*>(base + row_chunk_offset(row, ch)) = make_float4(r[4 * ch], r[4 * ch + 1], r[4 * ch + 2], r[4 * ch + 3]);
}

// issue D[128 x n] (+)= A[128 x 32] . B[n x 32]^T as four K=8 steps; A in TMEM columns [a_tmem, a_tmem + 32).  One thread.
// The four descriptors differ only in the start-address field (+256 B = +16 units per step), so the low word is
// built once and incremented: the issuer's instructions come out of the issue slots of the scan warp that
// shares its SM sub-partition.
__device__ __forceinline__ void mma_tf32_ts_desc(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void issue_tile_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_smem, uint32_t n) {
  const uint32_t idesc = make_idesc_tf32(128, n);
  const uint64_t d0 = make_smem_desc(b_smem);
  const uint32_t lo = (uint32_t)d0, hi = (uint32_t)(d0 >> 32);
#pragma unroll
  for (uint32_t s = 0; s < 4; ++s) mma_tf32_ts_desc(d_tmem, a_tmem + s * 8u, lo + s * 16u, hi, idesc, s);
}
// ---------------------------------------------------------------------------------------
// Warp-specialised pipeline of one CTA (kThreads = 768 threads = 6 warpgroups, 1 CTA per SM; launched
// at 80 registers per thread, then setmaxnreg redistributes the pool: shade 72, issuer warpgroup 24,
// scan warpgroup 168 -- its two 32-register TMEM loads in flight plus the packing do not fit in 80):
//   SHADE   4 groups x 128 threads (warps 0-15): each thread owns one path; per round it publishes
//           its lifted ray as row t of the group's A tile (+ its threshold P), waits for the group's
//           candidate bitmap, resolves the closest hit and shades.
//   ISSUER  warp 20 (warps 21-23 only donate their registers).  The whole warp runs the loop and waits
//           convergently, one elect.sync lane issues: for pass m = 0, 1, ... and group g = 0..3 IN THAT FIXED
//           ORDER it waits for the group's A tile and issues the MMAs of every sphere chunk into a ring of kRing
//           TMEM slots of kNC columns (3 x 128 columns; the other 128 columns hold the four A tiles).
//           Swept operand: it also streams the next slice of the sphere operand (issuer_role).
//   SCAN    warps 16-19 (warp 16+q owns TMEM lanes 32q..32q+31): consume the ring in the same fixed
//           order; thread t reads the f values of ray row t of the current group with
//           tcgen05.ld.32x32b.x32 (two loads in flight) and packs the 32 sign bits of every load into
//           one word of that ray's candidate BITMAP in shared memory (funnel shifts: no compare,
//           no branch, no divergence; bit 31 of word w = sphere 32w).
// Barriers (all mbarriers with deterministic phase parity; a parity wait only tells adjacent phases apart,
// so every barrier is waited on by a party that visits each of its phases in turn):
//   a_ready[g]   128 arrivals (shade threads of group g): the A tile of the round is stored / the bitmap
//                of the previous pass has been consumed
//   tmem_full[s] 1 arrival (tcgen05.commit): the MMAs of the chunk in ring slot s are complete
//   tmem_empty[s] 4 arrivals (scan warps): slot s has been read into registers
//   scan_done[g] 128 arrivals (scan threads): the bitmaps of group g for this pass are complete
//   b_full[i], b_free[i] (swept operand only): slice buffer i has landed / all MMAs reading it are complete
// Ring slots are numbered by a sequence number running across groups and passes (slot = seq % kRing,
// use = seq / kRing, parity = use & 1; both are kept incrementally in the loops).
// Termination: every shade warp stores "all my paths are done" for round r into done_flags[r & 7]
// before its a_ready arrive; EVERY role leaves at the first round r >= 2 for which all 16 flags of
// round r-2 are set.  All roles are within two rounds of each other and the flags of round r-2
// are ordered before anything a role does in round r by the barrier chain, so all roles read the
// same values and leave at the same round.  (fixed_rounds >= 0: exactly that many rounds, K2.)
#ifndef UECRT_TC_GROUPS
#define UECRT_TC_GROUPS 4          // shade groups of 128 paths
#endif
#ifndef UECRT_TC_SCAN_WGS
#define UECRT_TC_SCAN_WGS 1        // scan warpgroups; 2: they split the 128 columns of every chunk in halves (two scan warps per SM sub-partition)
#endif
#ifndef UECRT_TC_SHADE_REGS
#define UECRT_TC_SHADE_REGS 88
#endif
constexpr int kGroups = UECRT_TC_GROUPS;
constexpr int kGroupThreads = 128;
constexpr int kShadeThreads = kGroups * kGroupThreads;
constexpr int kScanGroups = UECRT_TC_SCAN_WGS;
constexpr int kScanThreads = kScanGroups * 128;
constexpr int kThreads = kShadeThreads + kScanThreads + 128; // + the issuer warpgroup (one working warp)
// Thread ranges of the roles.  Within an SM sub-partition the warp scheduler prefers the eligible warp with the HIGHEST
// warp id (B300_MICROARCH.md, multi-warp arbiter), so the scan warps -- the pipeline's pacing role -- come last; the issuer's
// barrier polling sits between them and the shade warps.  (Round 1 had the issuer last: its polling took issue slots from the
// scan warp of its sub-partition, which then ran every phase 1.3 - 2.7x slower than the other three and paced the ring.)
#ifndef UECRT_TC_SCAN_LAST
#define UECRT_TC_SCAN_LAST 1
#endif
#ifndef UECRT_TC_NAMED_BARRIERS
#define UECRT_TC_NAMED_BARRIERS 0   // 1: bar.sync / bar.arrive hand-offs instead of mbarriers (measured 5 % slower: 19.06 vs 18.03 ms)
#endif
// named barrier ids (0 is __syncthreads)
constexpr uint32_t kBarAReady = 1, kBarScanDone = 1 + UECRT_TC_GROUPS, kBarEmpty = 1 + 2 * UECRT_TC_GROUPS;
constexpr uint32_t kBarAReadyCount = 128 + 32;      // the group's shade threads arrive, the issuer warp waits
constexpr uint32_t kBarEmptyCount = 128 + 32;       // the scan warpgroup arrives, the issuer warp waits
constexpr uint32_t kBarScanDoneCount = 128 + 128;   // the scan warpgroup arrives, the group's shade threads wait
constexpr int kIssuerFirstThread = UECRT_TC_SCAN_LAST ? kShadeThreads : kShadeThreads + kScanThreads;
constexpr int kScanFirstThread = UECRT_TC_SCAN_LAST ? kShadeThreads + 128 : kShadeThreads;
__device__ __forceinline__ bool is_issuer_thread() { return (int)threadIdx.x >= kIssuerFirstThread && (int)threadIdx.x < kIssuerFirstThread + 128; }
__device__ __forceinline__ bool is_scan_thread() { return (int)threadIdx.x >= kScanFirstThread && (int)threadIdx.x < kScanFirstThread + kScanThreads; }
static_assert(1 + 2 * UECRT_TC_GROUPS + 3 <= 16, "named barrier ids");
static_assert(kThreads == 768 || kThreads == 896, "24 warps launch at 80 registers, 28 warps at 72 (7 warps per SM sub-partition x 72 x 32 <= 16 K)");
// setmaxnreg targets: the launch allocation (80 or 72 registers x kThreads) is redistributed between the roles
constexpr int kLaunchRegs = kThreads == 768 ? 80 : 72;
constexpr int kShadeRegs = UECRT_TC_SHADE_REGS, kIssuerRegs = 24;
constexpr int kScanRegsFit = ((kThreads * kLaunchRegs - kShadeThreads * kShadeRegs - 128 * kIssuerRegs) / kScanThreads) / 8 * 8;
constexpr int kScanRegs = kScanRegsFit > 128 ? 128 : kScanRegsFit;
constexpr int kNC = 128;                                     // spheres (TMEM columns) per ring slot
constexpr int kRing = 3;                                     // ring slots: TMEM columns [0, 384)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kATmemCol = kRing * kNC;                  // A tile of group g: TMEM columns [384 + 32g, 384 + 32g + 32)
#ifndef UECRT_TC_QUAD
#define UECRT_TC_QUAD 4            // spheres per candidate-bitmap bit: 1 (one bit per sphere) or 4 (one bit per quad, re-tested in FP32 by the shade thread)
#endif
constexpr int kQuad = UECRT_TC_QUAD;                         // spheres per candidate-bitmap bit (tc_collect.cuh)
static_assert(kQuad == 1 || kQuad == 4, "bitmap granularity");
constexpr int kWordsPerChunk = 128 / (32 * kQuad);           // bitmap words per ray and 128-sphere chunk (4 or 1)
constexpr int kMaxRows = kQuad == 1 ? 1024 : 1536;           // one bit per sphere: a 32-bit summary word per ray covers 32 bitmap words
constexpr uint32_t kHeaderBytes = 1024;

struct Header {               // first kHeaderBytes of dynamic shared memory
  uint64_t a_ready[kGroups];
  uint64_t scan_done[kGroups];
  uint64_t tmem_full[kRing];
  uint64_t tmem_empty[kRing];
  uint64_t b_loaded;          // the sphere operand has landed (cp.async.bulk)
  uint64_t b_full[2];         // swept operand: the pass in buffer i has landed (expect_tx + cp.async.bulk)
  uint64_t b_free[2];         // swept operand: every MMA that reads buffer i has completed (tcgen05.commit)
  uint32_t tmem_base;
  uint32_t pad_;
  uint8_t done_flags[8][16];  // [round & 7][shade warp] (kShadeThreads / 32 of the 16 are used)
};

// smem layout: [Header][B: nbuf x rows * 128 B][bitmap: rows/128 words x 512 rays]
// (nbuf = 1: the whole sphere operand is resident; nbuf = 2: it is swept through two buffers of `rows` rows, see issuer_role)
// One bitmap word per ray and 128-sphere chunk: bit (31 - k) = some sphere of QUAD k of the chunk (spheres 4k .. 4k+3)
// may be hit.  Word c of ray i lives at bitmap[c * 512 + i]: the 32 lanes of a warp touch 32 different banks
// whatever chunk each of them is at.
// Resident operand: TWO bitmap buffers (tile m of a group goes to buffer m & 1): a shade thread walks the bitmap of
// one of its two paths while the scan warps write the other's (kernels.cu, k_render_paths_tc).  Swept operand: one.
// Resident operand only: the shade side's tables follow the bitmaps, the quad table (64 B per quad, scene_pack.hpp
// pack_scene_quads) and the spheres' float4 {cx,cy,cz,r}: a flagged quad costs shared-memory loads, not an L1 / L2 round trip.
// One bit per sphere (kQuad == 1): bitmap word w of ray i covers spheres 32 w .. 32 w + 31 of the pass and the ray's SUMMARY
// word (one per bitmap buffer, in front of it) has bit (31 - w) set iff word w is non-zero, so the shade thread jumps to its
// next non-empty word.
struct Layout {
  uint32_t b, bitmap, bitmap_buf_bytes, quads, geom, total;
};
__host__ __device__ inline Layout make_layout(uint32_t b_rows, uint32_t nbuf = 1u) {
  Layout l;
  l.b = kHeaderBytes;
  l.bitmap = l.b + nbuf * b_rows * 128u;
  // per buffer: [summary: one word per ray and scan warpgroup][words: b_rows / (32 kQuad) per ray]
  l.bitmap_buf_bytes = ((uint32_t)kScanGroups + (b_rows / (uint32_t)kNC) * (uint32_t)kWordsPerChunk) * (uint32_t)kShadeThreads * 4u;
  l.quads = l.bitmap + (nbuf == 1u ? 2u : 1u) * l.bitmap_buf_bytes;
  l.geom = l.quads + (nbuf == 1u ? b_rows * 16u : 0u);
  l.total = l.geom + (nbuf == 1u ? b_rows * 16u : 0u);
  return l;
}

template <int R> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// bulk copy global -> shared in pieces, completing `bytes` on the mbarrier (one thread)
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  const uint32_t kPiece = 32768;
  for (uint32_t off = 0; off < bytes; off += kPiece) {
    const uint32_t sz = bytes - off < kPiece ? bytes - off : kPiece;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem + off),
                 "l"((const char*)src + off), "r"(sz), "r"(bar)
                 : "memory");
  }
}

// CTA-wide setup: barriers, TMEM, the bulk copy of the (resident) sphere operand; b_bytes = 0: the operand
// is swept, the issuer loads it pass by pass.  Every thread calls it.
__device__ __forceinline__ uint32_t cta_setup(unsigned char* smem_raw, const void* b_global, uint32_t b_bytes, uint32_t nchunks_per_tile) {

  Header* h = reinterpret_cast<Header*>(smem_raw);
  if (threadIdx.x == 0) {
    for (int g = 0; g < kGroups; ++g) {
      mbar_init(smem_addr(&h->a_ready[g]), kGroupThreads / 32);   // one arrival per shade warp
      mbar_init(smem_addr(&h->scan_done[g]), 4u * nchunks_per_tile);   // one arrival per scan warp and chunk of the tile
    }
    for (int s = 0; s < kRing; ++s) {
      mbar_init(smem_addr(&h->tmem_full[s]), 1);
      mbar_init(smem_addr(&h->tmem_empty[s]), 4);                 // the four warps of the scan warpgroup
    }
    mbar_init(smem_addr(&h->b_loaded), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_addr(&h->b_full[i]), 1); mbar_init(smem_addr(&h->b_free[i]), 1); }
    for (int i = 0; i < 8 * 16; ++i) (&h->done_flags[0][0])[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((threadIdx.x >> 5) == 0) tmem_alloc(smem_addr(&h->tmem_base), kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (threadIdx.x == 0 && b_bytes) bulk_load(smem_addr(smem_raw + kHeaderBytes), b_global, b_bytes, smem_addr(&h->b_loaded));
  return h->tmem_base;
}
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base) {
  fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem_base, kTmemCols);
}

__device__ __forceinline__ bool all_done_at(const Header* h, uint32_t r) {
  const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(&h->done_flags[r & 7u][0]);
  uint32_t all = 0x01010101u;
#pragma unroll
  for (int w = 0; w < kShadeThreads / 128; ++w) all &= f[w];     // four warp flags per word
  return all == 0x01010101u;
}
__device__ __forceinline__ bool round_continues(const Header* h, uint32_t r, int fixed_rounds) {
  if (fixed_rounds >= 0) return r < (uint32_t)fixed_rounds;
  return !(r >= 2u && all_done_at(h, r - 2u));
}

// The pipeline's unit of work is a PASS: every group's tile against `nchunks` chunks of the sphere operand.
// Resident operand (scenes up to kMaxRows rows): one pass per round.  SWEPT operand (larger scenes): a round is
// `npasses` passes over consecutive slices of the operand, which stream from L2 through two shared-memory
// buffers; the ray tiles stay in TMEM for the whole round and the shade threads fold every pass's
// candidates into their running closest hit.  Passes are numbered m = 0, 1, ... across rounds; every
// per-pass barrier parity below is a function of m only.  The termination vote is per round and read
// at a round's first pass.
struct PassIter {
  uint32_t m = 0, pass = 0, round = 0;
  __device__ __forceinline__ bool begin(const Header* h, int fixed_rounds) const { return pass != 0u || round_continues(h, round, fixed_rounds); }
  __device__ __forceinline__ void next(uint32_t npasses) { ++m; if (++pass == npasses) { pass = 0; ++round; } }
};

#ifndef UECRT_TC_ISSUER_LANES
#define UECRT_TC_ISSUER_LANES 32    // 32: the whole warp runs the loop convergently and one elected lane issues (2.5 - 5 % faster than 1: one elected lane running the whole role in a diverged warp)
#endif
// ISSUER: one thread.  SWEEP: it also streams the operand: the slice of pass m+1 is requested during pass m
// (behind its first tile) into buffer (m+1) & 1, once the MMAs of pass m-1 (the buffer's previous readers) have
// completed (b_free, a tcgen05.commit behind that pass's last MMA); pass m starts when its slice has landed (b_full).
template <bool SWEEP>
__device__ __forceinline__ void issuer_role(unsigned char* smem_raw, uint32_t tmem_base, int nchunks, uint32_t npasses,
                                            const unsigned char* b_global, int fixed_rounds,
                                            unsigned long long* prof_counters = nullptr) {
#ifdef UECRT_TC_PROFILE
  long long p_ready = 0, p_empty = 0, p_issue = 0, p_bfree = 0, p_bfull = 0, p_last = clock64();
  unsigned long long p_tiles = 0;
#define UECRT_TCI(acc) { const long long now_ = clock64(); acc += now_ - p_last; p_last = now_; }
#else
#define UECRT_TCI(acc)
#endif
  Header* h = reinterpret_cast<Header*>(smem_raw);
  const Layout lay = make_layout((uint32_t)nchunks * kNC, SWEEP ? 2u : 1u);
  const uint32_t b_smem0 = smem_addr(smem_raw + lay.b), pass_bytes = (uint32_t)nchunks * (kNC * 128u);
  uint32_t full0 = smem_addr(&h->tmem_full[0]), empty0 = smem_addr(&h->tmem_empty[0]), ready0 = smem_addr(&h->a_ready[0]);
  const uint32_t bfull0 = smem_addr(&h->b_full[0]), bfree0 = smem_addr(&h->b_free[0]);
  pin(full0); pin(empty0); pin(ready0);
  uint32_t load_pass = 0;                                   // the operand slice the next load fetches
  auto load_slice = [&](uint32_t m) {                       // into buffer m & 1, completing on b_full[m & 1]
    if (UECRT_TC_ISSUER_LANES == 1 || elect_one())
      bulk_load(b_smem0 + (m & 1u) * pass_bytes, b_global + (size_t)load_pass * pass_bytes, pass_bytes, bfull0 + 8u * (m & 1u));
    if (++load_pass == npasses) load_pass = 0;
  };
  if (!SWEEP) mbar_wait(smem_addr(&h->b_loaded), 0);
  else { load_slice(0); load_slice(1); }
  uint32_t seq = 0, slot = 0, parity = 1;                   // slot = seq % kRing; parity = that of the slot's PREVIOUS use, ((seq / kRing) - 1) & 1
  PassIter it;
  for (; it.begin(h, fixed_rounds); it.next(npasses)) {
    const uint32_t m = it.m;
    if (SWEEP) {
      UECRT_TCI(p_issue)
      mbar_wait(bfull0 + 8u * (m & 1u), (m >> 1) & 1u);
      UECRT_TCI(p_bfull)
    }
    const uint32_t b_smem = b_smem0 + (SWEEP ? (m & 1u) * pass_bytes : 0u);
    for (uint32_t g = 0; g < (uint32_t)kGroups; ++g) {
      UECRT_TCI(p_issue)
      if (UECRT_TC_NAMED_BARRIERS) nbar_sync(kBarAReady + g, kBarAReadyCount);   // the group's A tile of this pass is in TMEM
      else mbar_wait(ready0 + 8u * g, m & 1u);
      UECRT_TCI(p_ready)
      for (int j = 0; j < nchunks; ++j, ++seq, slot = slot == kRing - 1 ? 0u : slot + 1u, parity ^= (slot == 0u ? 1u : 0u)) {
#ifdef UECRT_TC_TRACE
        const bool tri_ = prof_counters && blockIdx.x == 0 && seq >= 3000u && seq < 3048u;
        if (tri_) prof_counters[128 + 20 * (seq - 3000u)] = clock64();
#endif
        if (seq >= (uint32_t)kRing) {                         // the scanners have read the previous contents
          if (UECRT_TC_NAMED_BARRIERS) nbar_sync(kBarEmpty + slot, kBarEmptyCount);
          else mbar_wait(empty0 + 8u * slot, parity);
        }
        UECRT_TCI(p_empty)
#ifdef UECRT_TC_TRACE
        if (tri_) prof_counters[128 + 20 * (seq - 3000u) + 1] = clock64();
#endif
        fence_after_sync();
        if (UECRT_TC_ISSUER_LANES == 1 || elect_one()) {
          issue_tile_ts(tmem_base + slot * kNC, tmem_base + kATmemCol + g * (uint32_t)kSlots, b_smem + (uint32_t)j * (kNC * 128u), kNC);
          mma_commit(full0 + 8u * slot);
        }
        if (UECRT_TC_ISSUER_LANES != 1) __syncwarp();
#ifdef UECRT_TC_TRACE
        if (tri_) prof_counters[128 + 20 * (seq - 3000u) + 2] = clock64();
#endif
        UECRT_TCI(p_issue)
      }
#ifdef UECRT_TC_PROFILE
      ++p_tiles;
#endif
      if (SWEEP && g == 0u && m >= 1u) {
        // behind the first tile, not at the pass boundary: by now the previous pass's MMAs have long completed,
        // so the wait is free, and the ring already holds this pass's first chunks while the copy is set up
        UECRT_TCI(p_issue)
        mbar_wait(bfree0 + 8u * ((m - 1u) & 1u), ((m - 1u) >> 1) & 1u);
        UECRT_TCI(p_bfree)
        load_slice(m + 1u);
        UECRT_TCI(p_bfull)
      }
    }
    if (SWEEP && (UECRT_TC_ISSUER_LANES == 1 || elect_one())) mma_commit(bfree0 + 8u * (m & 1u));
  }
  if (SWEEP) {
    // loads were requested for passes 0 .. max(1, m_end); those never run: let them land before the CTA goes away
    const uint32_t last = it.m > 1u ? it.m : 1u;
    for (uint32_t k = it.m; k <= last; ++k) mbar_wait(bfull0 + 8u * (k & 1u), (k >> 1) & 1u);
  }
#ifdef UECRT_TC_PROFILE
  if (prof_counters) {
    atomicAdd(prof_counters + 20, p_tiles);
    atomicAdd(prof_counters + 21, (unsigned long long)p_ready);
    atomicAdd(prof_counters + 22, (unsigned long long)p_empty);
    atomicAdd(prof_counters + 23, (unsigned long long)p_issue);
    atomicAdd(prof_counters + 40, (unsigned long long)p_bfree);
    atomicAdd(prof_counters + 41, (unsigned long long)p_bfull);
  }
#endif
}

// SHADE side of one thread
struct Shade {
  uint32_t a_ready, scan_done;       // the group's barriers: shared addresses of the mbarriers (ids with UECRT_TC_NAMED_BARRIERS)
  uint32_t a_taddr;                  // TMEM address of this thread's A row: its warp's lane quarter | the group's A columns
  const uint32_t* bitmap;            // this ray's SUMMARY word(s) of bitmap buffer 0; bitmap word w at bitmap[(kScanGroups + w) * kShadeThreads]
  int nwords;                        // bitmap words per pass; buffer 1 starts (kScanGroups + nwords) * kShadeThreads words further
#ifdef UECRT_TC_PROFILE
  unsigned long long prof_trips = 0, prof_steps = 0, prof_cands = 0;
#endif
#ifdef UECRT_TC_TRACE
  unsigned long long* trace = nullptr;
#endif
};
__device__ __forceinline__ Shade make_shade(unsigned char* smem_raw, uint32_t tmem_base, int nchunks, uint32_t nbuf = 1u) {
  Header* h = reinterpret_cast<Header*>(smem_raw);
  const Layout lay = make_layout((uint32_t)nchunks * kNC, nbuf);
  const uint32_t g = threadIdx.x / kGroupThreads;
  Shade s;
  s.a_ready = UECRT_TC_NAMED_BARRIERS ? kBarAReady + g : smem_addr(&h->a_ready[g]);
  s.scan_done = UECRT_TC_NAMED_BARRIERS ? kBarScanDone + g : smem_addr(&h->scan_done[g]);
  s.a_taddr = tmem_base + kATmemCol + g * (uint32_t)kSlots + (((threadIdx.x >> 5) & 3u) * 32u << 16);
  s.bitmap = reinterpret_cast<const uint32_t*>(smem_raw + lay.bitmap) + threadIdx.x;
  s.nwords = nchunks * kWordsPerChunk;
  return s;
}
#endif  // __CUDACC__

}  // namespace tc
}  // namespace uecrt
