// tc_probe.cu -- validates and sizes the tensor-core candidate scan (csrc/tc_scan.cuh) on a B200.
// Not part of the product path: evidence for DESIGN.md (operand layout, error constant, throughput).
//   T1 layout   : small-integer operands (exact in TF32) -> D must equal the integer product
//   T2 precision: lifted (ray, sphere) vectors of a final-scene-like world; error of the 3-way TF32
//                 split + TMEM accumulation against the exact dot product, in units of u = 2^-24
//                 relative to S = sum |a_i||b_i|
//   T3 speed    : 148 CTAs x 4 groups of 128 rays sweeping N spheres per round (MMA + TMEM readout
//                 + integer-max reduction), cycles per round
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I uecraytracing_b200/csrc -o tc_probe tc_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "tc_scan.cuh"

using namespace uecrt::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int kPNC = 128;          // TMEM columns (spheres) per MMA tile
constexpr int kPGroupThreads = 128;

struct ProbeArgs {
  const float* a_rows;       // [n_groups_total * 128][32] pre-filled K-slot rows (T1/T2) or null (T3: synthesised)
  const unsigned char* b;    // canonical layout, n_spheres/8 row groups of 1024 B
  float* d_out;              // [n_groups_total * 128][n_spheres] or null
  unsigned long long* sink;  // T3: keeps the reduction alive
  int n_spheres;             // multiple of kPNC
  int rounds;                // T3
  int mode;                  // 0: full, 1: MMA only (no readout), 2: readout only (no MMA)
  uint32_t lbo, sbo;         // descriptor fields under test
};

__device__ __forceinline__ uint64_t desc_ex(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void issue_tile_ex(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem, uint32_t n, uint32_t lbo, uint32_t sbo) {
  const uint32_t idesc = make_idesc_tf32(128, n);
#pragma unroll
  for (uint32_t s = 0; s < 4; ++s)
    mma_tf32(d_tmem, desc_ex(a_smem + s * 256u, lbo, sbo), desc_ex(b_smem + s * 256u, lbo, sbo), idesc, s);
}

// smem: [0,64) barriers + tmem ptr | B (n_spheres * 128 B) | A (groups * 16 KB)
template <int GROUPS>
__global__ void __launch_bounds__(GROUPS * kPGroupThreads, 1) k_probe(const ProbeArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // GROUPS barriers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 48);
  unsigned char* sB = smem + 1024;
  unsigned char* sA = sB + (size_t)p.n_spheres * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = tid / kPGroupThreads, tg = tid % kPGroupThreads;

  if (tid == 0) {
    for (int i = 0; i < GROUPS; ++i) mbar_init(smem_addr(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_addr(tmem_ptr), 512);
  // B: generic-proxy copy, then make it visible to the async proxy (the product uses cp.async.bulk)
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.b);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    const int n16 = p.n_spheres * 8;
    for (int i = tid; i < n16; i += blockDim.x) dst[i] = src[i];
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  unsigned char* myA = sA + (size_t)g * 128 * 128;
  const uint32_t a_smem = smem_addr(myA), b_smem = smem_addr(sB);
  const uint32_t bar = smem_addr(bars + g);
  const uint32_t d_tmem = tmem_base + (uint32_t)g * kPNC;
  const uint32_t my_taddr = d_tmem + ((uint32_t)(warp & 3) * 32u << 16);
  const bool elected = (tg == 0);
  const int nchunks = p.n_spheres / kPNC;
  const size_t grow = ((size_t)blockIdx.x * GROUPS + g) * 128 + tg;   // global ray row
  uint32_t phase = 0;
  unsigned long long acc = 0;

  const int rounds = p.a_rows ? 1 : p.rounds;
  for (int r = 0; r < rounds; ++r) {
    float row[kSlots];
    if (p.a_rows) {
#pragma unroll
      for (int i = 0; i < kSlots; ++i) row[i] = p.a_rows[grow * kSlots + i];
    } else {
      float a[kTerms];
#pragma unroll
      for (int i = 0; i < kTerms; ++i) a[i] = __uint_as_float(0x3f000000u + ((uint32_t)(tid * 977 + r * 131 + i * 7919) & 0x7fffffu));
      fill_row_a(a, 0.0f, row);
    }
    store_row(myA, tg, row);
    fence_proxy_async_smem();
    fence_before_sync();
    group_bar(1 + g, kPGroupThreads);
    if (p.mode != 2 && elected) {
      fence_after_sync();
      issue_tile_ex(d_tmem, a_smem, b_smem, kPNC, p.lbo, p.sbo);
      mma_commit(bar);
    }
    for (int j = 0; j < nchunks; ++j) {
      if (p.mode != 2) { mbar_wait(bar, phase); phase ^= 1u; }
      fence_after_sync();
      if (p.mode != 1) {
#pragma unroll
        for (int c = 0; c < kPNC / 32; ++c) {
          int v[32];
          tmem_ld32(my_taddr + c * 32, v);
          tmem_ld_wait();
          if (p.d_out) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p.d_out[grow * p.n_spheres + j * kPNC + c * 32 + i] = __int_as_float(v[i]);
          } else {
            int m = v[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) m = max(m, v[i]);
            if (m >= 0x7f000000) acc += (unsigned)m;
          }
        }
      }
      if (j + 1 < nchunks) {
        fence_before_sync();
        group_bar(1 + g, kPGroupThreads);
        if (p.mode != 2 && elected) {
          fence_after_sync();
          issue_tile_ex(d_tmem, a_smem, b_smem + (uint32_t)(j + 1) * (kPNC / 8) * kRowGroupBytes, kPNC, p.lbo, p.sbo);
          mma_commit(bar);
        }
      }
    }
  }
  if (p.sink && acc) atomicAdd(p.sink, acc);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- T4: tensor-pipe issue patterns.  One thread issues `count` MMAs (M=128, N=n, K=8 TF32, B from shared
// memory, A from shared memory or TMEM), either all accumulating into ONE D tile (a dependent chain) or
// round-robin over `nacc` D tiles, then commits and waits.  Cycles per MMA by clock64.
template <bool ELECT>
__global__ void __launch_bounds__(128, 1) k_mma_pattern(int n, int nacc, int count, int a_in_tmem, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 16);
  unsigned char* sA = smem + 1024;            // 128 rows x 32 slots
  unsigned char* sB = sA + 16384;             // 256 rows x 32 slots
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(smem_addr(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) tmem_alloc(smem_addr(tmem_ptr), 512);
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(sA)[i] = 0.001f * (float)(i % 97);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  if (a_in_tmem) {      // A rows into TMEM columns [480, 512)
    float row[32];
    for (int i = 0; i < 32; ++i) row[i] = 0.001f * (float)((tid * 32 + i) % 97);
    tmem_st32(tmem_base + 480 + (((uint32_t)tid >> 5) * 32u << 16), row);
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  bool issuer;
  if (ELECT) {
    issuer = false;
    if (__shfl_sync(0xffffffffu, tid >> 5, 0) == 0) issuer = elect_one();
  } else {
    issuer = (tid == 0);
  }
  if (issuer) {
    const uint32_t idesc = make_idesc_tf32(128, n);
    const uint32_t a_smem = smem_addr(sA), b_smem = smem_addr(sB);
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t d = tmem_base + (uint32_t)(i % nacc) * (uint32_t)n;
      const uint32_t ks = (uint32_t)(i & 3);
      if (a_in_tmem) mma_tf32_ts(d, tmem_base + 480 + ks * 8u, make_smem_desc(b_smem + ks * 256u), idesc, 1);
      else mma_tf32(d, make_smem_desc(a_smem + ks * 256u), make_smem_desc(b_smem + ks * 256u), idesc, 1);
    }
    const long long t1 = clock64();
    mma_commit(smem_addr(bar));
    mbar_wait(smem_addr(bar), 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

static void pack_b(const std::vector<float>& rows, int n, std::vector<unsigned char>& out) {
  out.assign((size_t)n * 128, 0);
  for (int r = 0; r < n; ++r)
    for (int ch = 0; ch < 8; ++ch)
      memcpy(&out[row_chunk_offset(r, ch)], &rows[(size_t)r * kSlots + ch * 4], 16);
}

template <int GROUPS>
static float run(const ProbeArgs& a, int blocks) {
  const size_t smem = 1024 + (size_t)a.n_spheres * 128 + (size_t)GROUPS * 16384;
  CK(cudaFuncSetAttribute(k_probe<GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k_probe<GROUPS><<<blocks, GROUPS * kPGroupThreads, smem>>>(a);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

static double urand() { return (double)rand() / ((double)RAND_MAX + 1.0); }

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 7;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);

  // ------------------------------------------------------------------ T1 layout
  if (which & 1) {
    const int N = 256, M = 128;
    std::vector<float> A((size_t)M * kSlots), B((size_t)N * kSlots);
    for (int r = 0; r < M; ++r) for (int k = 0; k < kSlots; ++k) A[(size_t)r * kSlots + k] = (float)(((r * 7 + k * 3) % 5) - 2);
    for (int n = 0; n < N; ++n) for (int k = 0; k < kSlots; ++k) B[(size_t)n * kSlots + k] = (float)(((n * 5 + k * 11) % 7) - 3);
    std::vector<unsigned char> Bp; pack_b(B, N, Bp);
    float* dA; unsigned char* dB; float* dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bp.size())); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
    const uint32_t variants[2][2] = {{kLBO, kSBO}, {kSBO, kLBO}};
    for (int v = 0; v < 2; ++v) {
      CK(cudaMemset(dD, 0xff, (size_t)M * N * 4));
      ProbeArgs a{dA, dB, dD, nullptr, N, 1, 0, variants[v][0], variants[v][1]};
      run<1>(a, 1);
      std::vector<float> D((size_t)M * N);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      long bad = 0;
      for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
        double e = 0; for (int k = 0; k < kSlots; ++k) e += (double)A[(size_t)r * kSlots + k] * B[(size_t)n * kSlots + k];
        if (D[(size_t)r * N + n] != (float)e) ++bad;
      }
      printf("T1 layout  LBO=%u SBO=%u : %ld / %d mismatches\n", variants[v][0], variants[v][1], bad, M * N);
      if (bad) {
        for (int r = 0; r < 3; ++r) { printf("   row %d got:", r); for (int n = 0; n < 6; ++n) printf(" %g", D[(size_t)r * N + n]); printf("\n   row %d exp:", r);
          for (int n = 0; n < 6; ++n) { double e = 0; for (int k = 0; k < kSlots; ++k) e += (double)A[(size_t)r * kSlots + k] * B[(size_t)n * kSlots + k]; printf(" %g", e); } printf("\n"); }
      }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  // ------------------------------------------------------------------ T2 precision
  if (which & 2) {
    srand(12345);
    const int N = 512, GROUPS_TOTAL = 256, M = GROUPS_TOTAL * 128;
    // spheres: ground r=1000, 3 big r=1, the rest r=0.2 on the [-11,11]^2 grid (RTIOW final-scene-like)
    std::vector<float> cx(N), cy(N), cz(N), rr(N);
    for (int i = 0; i < N; ++i) {
      if (i == 0) { cx[i] = 0; cy[i] = -1000; cz[i] = 0; rr[i] = 1000; }
      else if (i < 4) { cx[i] = (float)(-4 + 4 * (i - 1)); cy[i] = 1; cz[i] = 0; rr[i] = 1; }
      else { cx[i] = (float)(-11 + 22 * urand()); cy[i] = 0.2f; cz[i] = (float)(-11 + 22 * urand()); rr[i] = 0.2f; }
    }
    std::vector<float> brow((size_t)N * kSlots), bvec((size_t)N * kTerms);
    for (int i = 0; i < N; ++i) {
      const double x = cx[i], y = cy[i], z = cz[i], r = rr[i];
      float b[kTerms] = {(float)(x * x), (float)(y * y), (float)(z * z), (float)(x * y), (float)(x * z), (float)(y * z), cx[i], cy[i], cz[i], (float)(-(x * x + y * y + z * z - r * r))};
      float row[kSlots]; fill_row_b(b, row);
      for (int k = 0; k < kSlots; ++k) brow[(size_t)i * kSlots + k] = row[k];
      for (int k = 0; k < kTerms; ++k) bvec[(size_t)i * kTerms + k] = b[k];
    }
    std::vector<unsigned char> Bp; pack_b(brow, N, Bp);
    // rays: half from the camera (13,2,3) towards the scene, half from random points near the ground / spheres
    std::vector<float> arow((size_t)M * kSlots), avec((size_t)M * kTerms), osq(M);
    for (int i = 0; i < M; ++i) {
      float ox, oy, oz;
      if (i & 1) { ox = 13; oy = 2; oz = 3; }
      else { ox = (float)(-60 + 120 * urand()); oy = (float)(urand() < 0.5 ? 0.0 : 0.4 * urand()); oz = (float)(-60 + 120 * urand()); }
      double tx = (i & 1) ? (-11 + 22 * urand()) - ox : urand() * 2 - 1, ty = (i & 1) ? (0.2 - oy) + (urand() - 0.5) : urand() * 2 - 1,
             tz = (i & 1) ? (-11 + 22 * urand()) - oz : urand() * 2 - 1;
      const double il = 1.0 / sqrt(tx * tx + ty * ty + tz * tz);
      const float dx = (float)(tx * il), dy = (float)(ty * il), dz = (float)(tz * il);
      const float od = fmaf(oz, dz, fmaf(oy, dy, ox * dx));
      const float qx = -2.0f * fmaf(od, dx, -ox), qy = -2.0f * fmaf(od, dy, -oy), qz = -2.0f * fmaf(od, dz, -oz);
      float a[kTerms] = {dx * dx, dy * dy, dz * dz, 2.0f * dx * dy, 2.0f * dx * dz, 2.0f * dy * dz, qx, qy, qz, 1.0f};
      float row[kSlots]; fill_row_a(a, 0.0f, row);
      for (int k = 0; k < kSlots; ++k) arow[(size_t)i * kSlots + k] = row[k];
      for (int k = 0; k < kTerms; ++k) avec[(size_t)i * kTerms + k] = a[k];
      osq[i] = ox * ox + oy * oy + oz * oz;
    }
    float* dA; unsigned char* dB; float* dD;
    CK(cudaMalloc(&dA, arow.size() * 4)); CK(cudaMalloc(&dB, Bp.size())); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, arow.data(), arow.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
    ProbeArgs a{dA, dB, dD, nullptr, N, 1, 0, kLBO, kSBO};
    run<4>(a, GROUPS_TOTAL / 4);
    std::vector<float> D((size_t)M * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    const double u = ldexp(1.0, -24);
    double worstS = 0, worstSep = 0, worstFp32 = 0; long hist[8] = {0};
    for (int i = 0; i < M; ++i)
      for (int n = 0; n < N; ++n) {
        double e = 0, S = 0; float f32 = 0;
        for (int k = 0; k < kTerms; ++k) {
          const double t = (double)avec[(size_t)i * kTerms + k] * bvec[(size_t)n * kTerms + k];
          e += t; S += fabs(t);
          f32 = fmaf(avec[(size_t)i * kTerms + k], bvec[(size_t)n * kTerms + k], f32);
        }
        const double err = fabs((double)D[(size_t)i * N + n] - e);
        const double c2 = (double)cx[n] * cx[n] + (double)cy[n] * cy[n] + (double)cz[n] * cz[n];
        const double sep = osq[i] + 3 * c2 + (double)rr[n] * rr[n];     // separable majorant of S
        const double kS = err / (u * S), kSep = err / (u * sep);
        if (kS > worstS) worstS = kS;
        if (kSep > worstSep) worstSep = kSep;
        const double k32 = fabs((double)f32 - e) / (u * S);
        if (k32 > worstFp32) worstFp32 = k32;
        int bin = kS < 1 ? 0 : kS < 2 ? 1 : kS < 4 ? 2 : kS < 8 ? 3 : kS < 16 ? 4 : kS < 32 ? 5 : kS < 64 ? 6 : 7;
        ++hist[bin];
      }
    printf("T2 precision over %d x %d pairs: max |f_tc - exact| = %.2f u*S (S = sum|a_i b_i|), %.2f u*(|o|^2+3|c|^2+r^2); plain FP32 fma chain: %.2f u*S\n",
           M, N, worstS, worstSep, worstFp32);
    printf("   histogram of err/(u*S): <1:%ld <2:%ld <4:%ld <8:%ld <16:%ld <32:%ld <64:%ld >=64:%ld\n", hist[0], hist[1], hist[2], hist[3], hist[4], hist[5], hist[6], hist[7]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  // ------------------------------------------------------------------ T5 accumulation error of the MMA itself, adversarial operands
  // Operands that are already TF32-exact (no split error): what remains is how tcgen05.mma.kind::tf32 sums 32 exact
  // products (four chained K = 8 steps into an FP32 accumulator in TMEM).  Families of rows / columns:
  //   0 same sign, equal magnitude, full 11-bit mantissas     1 one dominant slot (2^12 x the rest)
  //   2 alternating signs, equal magnitude (sum ~ 0, S large) 3 geometric spread 2^-k      4 random signs and exponents
  // err / (u S), u = 2^-24, S = sum |a_k b_k|, maximum per (row family, column family).
  if (which & 16) {
    srand(777);
    const int N = 640, GROUPS_TOTAL = 320, M = GROUPS_TOTAL * 128, F = 5;
    auto tf = [](double mant01, int e, int sign) {   // TF32-exact: 1.xxxxxxxxxx (10 fraction bits) x 2^e
      const double m = 1.0 + floor(mant01 * 1024.0) / 1024.0;
      return (float)(sign * ldexp(m, e));
    };
    auto fill = [&](float* row, int fam) {
      for (int k = 0; k < kSlots; ++k) {
        switch (fam) {
          case 0: row[k] = tf(urand(), 0, 1); break;
          case 1: row[k] = tf(urand(), k == 5 ? 6 : 0, 1); break;
          case 2: row[k] = tf(urand(), 0, (k & 1) ? -1 : 1); break;
          case 3: row[k] = tf(urand(), -(k / 2), 1); break;
          default: row[k] = tf(urand(), (int)(urand() * 12) - 6, urand() < 0.5 ? -1 : 1); break;
        }
      }
    };
    std::vector<float> arow((size_t)M * kSlots), brow((size_t)N * kSlots);
    for (int i = 0; i < M; ++i) fill(&arow[(size_t)i * kSlots], i % F);
    for (int n = 0; n < N; ++n) fill(&brow[(size_t)n * kSlots], n % F);
    std::vector<unsigned char> Bp; pack_b(brow, N, Bp);
    float* dA; unsigned char* dB; float* dD;
    CK(cudaMalloc(&dA, arow.size() * 4)); CK(cudaMalloc(&dB, Bp.size())); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, arow.data(), arow.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
    ProbeArgs a{dA, dB, dD, nullptr, N, 1, 0, kLBO, kSBO};
    run<4>(a, GROUPS_TOTAL / 4);
    std::vector<float> D((size_t)M * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    const double u = ldexp(1.0, -24);
    double worst[F][F] = {}, worst_rel_sum[F][F] = {};
    long above = 0, below = 0;   // sign of the error: truncation toward zero / -inf shows as a bias
    for (int i = 0; i < M; ++i)
      for (int n = 0; n < N; ++n) {
        double e = 0, S = 0;
        for (int k = 0; k < kSlots; ++k) { const double t = (double)arow[(size_t)i * kSlots + k] * brow[(size_t)n * kSlots + k]; e += t; S += fabs(t); }
        const double d = (double)D[(size_t)i * N + n] - e;
        const double kS = fabs(d) / (u * S);
        if (kS > worst[i % F][n % F]) worst[i % F][n % F] = kS;
        const double kR = fabs(d) / (u * fmax(fabs(e), 1e-300));
        if (S < 1.0001 * fabs(e) && kR > worst_rel_sum[i % F][n % F]) worst_rel_sum[i % F][n % F] = kR;
        if (d > 0) ++above; else if (d < 0) ++below;
      }
    printf("T5 accumulate error, TF32-exact operands, %d x %d pairs; max err / (u S) by (row family, column family):\n", M, N);
    double all = 0;
    for (int r = 0; r < F; ++r) { printf("   row family %d:", r); for (int c = 0; c < F; ++c) { printf(" %6.2f", worst[r][c]); if (worst[r][c] > all) all = worst[r][c]; } printf("\n"); }
    printf("   overall max %.2f u*S; same-sign sums (S = |sum|): max err / (u |sum|) = %.2f; result above exact %ld, below %ld (a truncating adder is always below for positive sums)\n",
           all, worst_rel_sum[0][0], above, below);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  // ------------------------------------------------------------------ T3 speed
  if (which & 4) {
    const int N = 512, rounds = 2000, blocks = prop.multiProcessorCount;
    std::vector<float> brow((size_t)N * kSlots);
    for (size_t i = 0; i < brow.size(); ++i) brow[i] = (float)((int)(i % 13) - 6) * 0.125f;
    std::vector<unsigned char> Bp; pack_b(brow, N, Bp);
    unsigned char* dB; unsigned long long* sink;
    CK(cudaMalloc(&dB, Bp.size())); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
    const char* names[3] = {"MMA + readout", "MMA only", "readout only"};
    for (int mode = 0; mode < 3; ++mode) {
      ProbeArgs a{nullptr, dB, nullptr, sink, N, rounds, mode, kLBO, kSBO};
      run<4>(a, blocks);                       // warm-up
      const float ms = run<4>(a, blocks);
      const double cyc = ms * 1e-3 * 1.965e9 / rounds;
      printf("T3 speed   %-14s: %.3f ms, %.0f cycles per round of 4 groups x 128 rays x %d spheres per SM = %.2f cycles per (warp, sphere) per SMSP; %.2f Gtests/s chip\n",
             names[mode], ms, cyc, N, cyc / (4.0 * N), (double)blocks * 512.0 * N * rounds / (ms * 1e-3) * 1e-9);
    }
    cudaFree(dB); cudaFree(sink);
  }
  // ------------------------------------------------------------------ T4 MMA issue patterns
  if (which & 8) {
    long long* dout; CK(cudaMalloc(&dout, 16));
    const size_t smem = 1024 + 16384 + 32768;
    CK(cudaFuncSetAttribute(k_mma_pattern<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_mma_pattern<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int count = 4096;
    for (int use_elect = 0; use_elect < 2; ++use_elect)
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
      for (int n : {64, 128, 256})
        for (int nacc : {1, 2, 3}) {
          if (n * nacc > 480) continue;
          long long h[2];
          for (int rep = 0; rep < 2; ++rep) {
            if (use_elect) k_mma_pattern<true><<<1, 128, smem>>>(n, nacc, count, a_tmem, dout);
            else k_mma_pattern<false><<<1, 128, smem>>>(n, nacc, count, a_tmem, dout);
            CK(cudaDeviceSynchronize());
          }
          CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
          printf("T4 mma     %s, A in %s, N=%3d, %d accumulator(s): issue %.1f cycles/MMA, issue+execute %.1f cycles/MMA\n", use_elect ? "elect.sync" : "if (tid==0)", a_tmem ? "TMEM" : "smem",
                 n, nacc, (double)h[0] / count, (double)h[1] / count);
        }
    cudaFree(dout);
  }
  return 0;
}
