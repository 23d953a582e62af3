// tc_scan.cuh -- tensor-core (tcgen05 / TMEM) form of the closest-hit candidate scan.
//
// The scan of closest_hit.cuh evaluates, for every (ray, sphere) pair,
//     f = (c.d)^2 + q.c - k          (candidate iff f >= P; sphere.hpp:27-32 with |d| = 1)
// which is BILINEAR in a lifted ray vector and a lifted sphere vector:
//     f = sum_{i<=j} (d_i d_j)(c_i c_j) + q.c + 1*(-k)
//       = a . b,   a = [dx^2, dy^2, dz^2, 2dxdy, 2dxdz, 2dydz, qx, qy, qz, 1]      (per ray)
//                  b = [cx^2, cy^2, cz^2, cxcy,  cxcz,  cycz,  cx, cy, cz, -k]     (per sphere)
// so the all-pairs test of 128 rays against N spheres is a 128 x N x 10 matrix product.  It runs on
// the 5th-generation tensor cores as tcgen05.mma.kind::tf32 with FP32 accumulation in TMEM.  TF32
// keeps 11 significand bits, so every operand is split a = a_hi + a_lo (each exactly representable
// in TF32) and the product is taken as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi: 3 x 10 = 30 of the 32
// K slots of four K=8 MMA steps.  All products of TF32 values are exact in FP32; what is lost is
// a_lo*b_lo (<= 2^-22 |a||b|), the truncation of the lo parts (<= 2^-21 each) and the accumulator's
// own rounding.  The bound on the sum of those is applied exactly like the FP32 scan's bound
// (scene_pack.hpp): the sphere part raises -k, the ray part lowers P, so {f >= P} stays a SUPERSET
// of the reference's "discriminant >= 0" and the resolve phase (closest_hit.cuh) is unchanged.
//
// Shared-memory operand layout (K-major, no swizzle; the UMMA "interleaved" canonical layout):
//   a row (ray or sphere) has 32 FP32 slots = 8 chunks of 16 B; rows are grouped by 8;
//   byte offset of (row, chunk) = (row / 8) * 1024 + chunk * 128 + (row % 8) * 16
//   -> core matrix = 8 rows x 16 B contiguous (128 B); LBO (next core matrix along K) = 128 B;
//      SBO (next 8-row group) = 1024 B.  One MMA step consumes 8 slots = 2 chunks.
// TMEM: D tile = 128 lanes (rays) x NC columns (spheres) of FP32; thread t of warp w reads lane
// 32*(w%4)+t with tcgen05.ld.32x32b, i.e. each thread reads the f values of ITS OWN ray.
#pragma once
#include <cuda_runtime.h>
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

// hi = v rounded to nearest TF32 (11 significant bits), lo = (v - hi) truncated to TF32.
// v - hi is exact in FP32 (Sterbenz-like: hi is v with the low bits rounded away).
__host__ __device__ inline void split_tf32(float v, float& hi, float& lo) {
#ifdef __CUDA_ARCH__
  const uint32_t b = __float_as_uint(v);
  hi = __uint_as_float((b + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(v - hi) & 0xffffe000u);
#else
  union { float f; uint32_t u; } x, y;
  x.f = v; x.u = (x.u + 0x1000u) & 0xffffe000u; hi = x.f;
  y.f = v - hi; y.u &= 0xffffe000u; lo = y.f;
#endif
}

// K-slot order of one row.  A (ray) side: [hi(10) | hi(10) | lo(10) | 0 0];  B (sphere) side:
// [hi(10) | lo(10) | hi(10) | 0 0]  ->  sum over slots = hi.hi + hi.lo + lo.hi.
__host__ __device__ inline void fill_row_a(const float (&a)[kTerms], float (&row)[kSlots]) {
  for (int i = 0; i < kTerms; ++i) {
    float h, l;
    split_tf32(a[i], h, l);
    row[i] = h; row[kTerms + i] = h; row[2 * kTerms + i] = l;
  }
  row[30] = 0.f; row[31] = 0.f;
}
__host__ __device__ inline void fill_row_b(const float (&b)[kTerms], float (&row)[kSlots]) {
  for (int i = 0; i < kTerms; ++i) {
    float h, l;
    split_tf32(b[i], h, l);
    row[i] = h; row[kTerms + i] = l; row[2 * kTerms + i] = h;
  }
  row[30] = 0.f; row[31] = 0.f;
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

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all tcgen05 operations issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
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

// write one 32-slot row (8 x 16 B) into the canonical layout; `base` = shared address of the
// operand tile, row = row within the tile
__device__ __forceinline__ void store_row(unsigned char* base, uint32_t row, const float (&r)[kSlots]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    *reinterpret_cast<float4*>(base + row_chunk_offset(row, ch)) = make_float4(r[4 * ch], r[4 * ch + 1], r[4 * ch + 2], r[4 * ch + 3]);
}

// issue D[128 x n] (+)= A[128 x 32] . B[n x 32]^T as four K=8 steps.  One thread.
__device__ __forceinline__ void issue_tile(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem, uint32_t n) {
  const uint32_t idesc = make_idesc_tf32(128, n);
#pragma unroll
  for (uint32_t s = 0; s < 4; ++s)
    mma_tf32(d_tmem, make_smem_desc(a_smem + s * 256u), make_smem_desc(b_smem + s * 256u), idesc, s);
}
// ---------------------------------------------------------------------------------------
// Producer / consumer pipeline of one CTA.
//   kGroups groups of 128 CONSUMER threads (one ray each; thread t of a group = TMEM lane t = row t
//   of the group's A tile) + one ISSUER warp per group (lane 0 issues the group's MMAs).
//   TMEM: group g owns 2 buffers of kNC columns at column (2g + b) * kNC (4 x 2 x 64 = 512).
//   Barriers per group: a_ready (128 arrivals: every consumer stored its A row), full[b] (tcgen05.commit:
//   the MMAs of one chunk into buffer b are complete), empty[b] (4 arrivals: every consumer warp has
//   read buffer b).  Chunks are numbered by a sequence number that runs across rounds; chunk s uses
//   buffer s & 1 and is the (s >> 1)-th use of it, which gives every barrier its phase parity.
constexpr int kGroups = 4;
constexpr int kGroupThreads = 128;
constexpr int kConsumers = kGroups * kGroupThreads;          // 512
constexpr int kThreads = kConsumers + 32 * kGroups;          // 640: + one issuer warp per group
constexpr int kNC = 64;                                      // spheres (TMEM columns) per chunk
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kATileBytes = 128u * kSlots * 4u;         // 16 KB per group
constexpr uint32_t kHeaderBytes = 1024;

struct GroupBars {            // 64 B
  uint64_t a_ready, full[2], empty[2];
  uint32_t stop;
  uint32_t pad[5];
};
struct Header {               // first kHeaderBytes of dynamic shared memory
  GroupBars g[kGroups];
  uint64_t b_loaded;          // the sphere operand has landed (cp.async.bulk)
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// named-barrier AND-reduction over the `count` threads that use barrier `id`
__device__ __forceinline__ bool group_all(uint32_t id, uint32_t count, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred pi, po;\n\tsetp.ne.u32 pi, %3, 0;\n\t"
      "bar.red.and.pred po, %1, %2, pi;\n\tselp.u32 %0, 1, 0, po;\n\t}"
      : "=r"(out)
      : "r"(id), "r"(count), "r"((uint32_t)pred)
      : "memory");
  return out != 0;
}

// smem layout of the tensor-scan kernels: [Header][B rows * 128][kGroups A tiles][lists]
__host__ __device__ inline size_t smem_bytes(uint32_t b_rows, uint32_t list_bytes) {
  return kHeaderBytes + (size_t)b_rows * 128 + (size_t)kGroups * kATileBytes + list_bytes;
}

// CTA-wide setup: barriers, TMEM, the bulk copy of the sphere operand.  Every thread calls it.
__device__ __forceinline__ uint32_t cta_setup(unsigned char* smem_raw, const void* b_global, uint32_t b_bytes) {
  Header* h = reinterpret_cast<Header*>(smem_raw);
  if (threadIdx.x == 0) {
    for (int g = 0; g < kGroups; ++g) {
      mbar_init(smem_addr(&h->g[g].a_ready), kGroupThreads);
      mbar_init(smem_addr(&h->g[g].full[0]), 1); mbar_init(smem_addr(&h->g[g].full[1]), 1);
      mbar_init(smem_addr(&h->g[g].empty[0]), 4); mbar_init(smem_addr(&h->g[g].empty[1]), 4);
      h->g[g].stop = 0;
    }
    mbar_init(smem_addr(&h->b_loaded), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((threadIdx.x >> 5) == 0) tmem_alloc(smem_addr(&h->tmem_base), kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (threadIdx.x == 0) {
    const uint32_t bar = smem_addr(&h->b_loaded);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_bytes) : "memory");
    const uint32_t kChunk = 32768;
    for (uint32_t off = 0; off < b_bytes; off += kChunk) {
      const uint32_t sz = b_bytes - off < kChunk ? b_bytes - off : kChunk;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr(smem_raw + kHeaderBytes + off)),
                   "l"((const char*)b_global + off), "r"(sz), "r"(bar)
                   : "memory");
    }
  }
  return h->tmem_base;
}
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base) {
  fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// ISSUER: lane 0 of issuer warp `g`.  Runs until the group's consumers raise `stop`.
__device__ __forceinline__ void issuer_loop(unsigned char* smem_raw, int g, uint32_t tmem_base, int nchunks) {
  Header* h = reinterpret_cast<Header*>(smem_raw);
  GroupBars* gb = &h->g[g];
  const uint32_t a_ready = smem_addr(&gb->a_ready), full0 = smem_addr(&gb->full[0]), empty0 = smem_addr(&gb->empty[0]);
  const uint32_t b_smem = smem_addr(smem_raw + kHeaderBytes);
  const uint32_t a_smem = b_smem + (uint32_t)nchunks * (kNC * 128u) + (uint32_t)g * kATileBytes;
  const uint32_t d_tmem = tmem_base + (uint32_t)g * 2u * kNC;
  const volatile uint32_t* stop = &gb->stop;
  mbar_wait(smem_addr(&h->b_loaded), 0);
  uint32_t seq = 0;
  for (uint32_t round = 0;; ++round) {
    mbar_wait(a_ready, round & 1u);
    if (*stop) break;
    for (int j = 0; j < nchunks; ++j, ++seq) {
      const uint32_t b = seq & 1u, use = seq >> 1;
      if (use >= 1u) mbar_wait(empty0 + 8u * b, (use - 1u) & 1u);      // the consumers have read the previous contents
      fence_after_sync();
      issue_tile(d_tmem + b * kNC, a_smem, b_smem + (uint32_t)j * (kNC * 128u), kNC);
      mma_commit(full0 + 8u * b);
    }
  }
}

// CONSUMER state of one thread
struct Consumer {
  uint32_t a_ready, full0, empty0;   // shared addresses of the group's barriers
  unsigned char* a_tile;             // the group's A tile
  uint32_t taddr;                    // TMEM address: this warp's lane quarter | the group's first column
  uint32_t row;                      // row in the A tile == TMEM lane
  uint32_t seq;                      // chunks consumed so far
  volatile uint32_t* stop;
};
__device__ __forceinline__ Consumer make_consumer(unsigned char* smem_raw, uint32_t tmem_base, int nchunks) {
  Header* h = reinterpret_cast<Header*>(smem_raw);
  const int g = threadIdx.x / kGroupThreads;
  Consumer c;
  c.a_ready = smem_addr(&h->g[g].a_ready); c.full0 = smem_addr(&h->g[g].full[0]); c.empty0 = smem_addr(&h->g[g].empty[0]);
  c.a_tile = smem_raw + kHeaderBytes + (size_t)nchunks * (kNC * 128u) + (size_t)g * kATileBytes;
  c.row = threadIdx.x % kGroupThreads;
  c.taddr = tmem_base + (uint32_t)g * 2u * kNC + ((((uint32_t)threadIdx.x >> 5) & 3u) * 32u << 16);
  c.seq = 0;
  c.stop = &h->g[g].stop;
  return c;
}
// the group has no more rays: tell its issuer to leave (every consumer of the group calls this once)
__device__ __forceinline__ void consumer_stop(Consumer& c) {
  if (c.row == 0) *c.stop = 1u;
  mbar_arrive(c.a_ready);
}
#endif  // __CUDACC__

}  // namespace tc
}  // namespace uecrt
