// Pipe-throughput microbenchmarks that size the K1 scan loop (B200, sm_100a).
// Not part of the product path: evidence for DESIGN.md's instruction budget.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float lo(unsigned long long v) {
  float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b;
}

// ---- 1. scalar FFMA, 8 independent chains, 3 register operands
__global__ void k_ffma(float* out, int iters, float m, float c) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m), "f"(c));
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 2. packed FFMA2
__global__ void k_ffma2(float* out, int iters, float m, float c) {
  unsigned long long a[8], mm = pk(m, m), cc = pk(c, c);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = pk(threadIdx.x * 1e-3f + i, 1.f + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(mm), "l"(cc));
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += lo(a[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 3. FP64 DFMA
__global__ void k_dfma(float* out, int iters, double m, double c) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
// ---- 4. MUFU sqrt.approx
__global__ void k_mufu(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i + 2.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 5. the scan-loop shaped mix: per sphere pair 2 broadcast LDS.128 + 8 packed FP ops
//         + sign-bit cull (LOP3s + ISETP) + a never-taken branch.  MODE 0: packed, MODE 1: scalar
template <int MODE>
__global__ void k_mix(float* out, int iters, int npairs, float fx, float fy, float fz) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < npairs * 8; i += blockDim.x) sm[i] = -1.0f - (i % 97) * 0.37f;
  __syncthreads();
  float t = threadIdx.x * 1e-4f;
  float dx = fx + t, dy = fy, dz = fz, ox = 2 * fx, oy = 2 * fy + t, oz = 2 * fz, od = fx * fy, oo = -1e9f;
  unsigned long long dx2 = pk(dx, dx), dy2 = pk(dy, dy), dz2 = pk(dz, dz), ox2 = pk(ox, ox), oy2 = pk(oy, oy),
                     oz2 = pk(oz, oz), od2 = pk(od, od), oo2 = pk(oo, oo);
  unsigned hitcount = 0;
  for (int it = 0; it < iters; ++it) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(sm);
#pragma unroll 4
    for (int j = 0; j < npairs; ++j) {
      ulonglong2 A = p[2 * j], B = p[2 * j + 1];   // {ncx01, ncy01}, {ncz01, k01}
      if (MODE == 0) {
        unsigned long long hb, cc, disc;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(hb) : "l"(A.x), "l"(dx2), "l"(od2));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hb) : "l"(A.y), "l"(dy2));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hb) : "l"(B.x), "l"(dz2));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(cc) : "l"(A.x), "l"(ox2), "l"(oo2));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(A.y), "l"(oy2));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(B.x), "l"(oz2));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(cc) : "l"(B.y));
        asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(disc) : "l"(hb));
        asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(disc) : "l"(cc));
        unsigned d0 = (unsigned)disc, d1 = (unsigned)(disc >> 32), h0 = (unsigned)hb, h1 = (unsigned)(hb >> 32),
                 c0 = (unsigned)cc, c1 = (unsigned)(cc >> 32);
        unsigned w0 = ~d0 & (h0 | c0), w1 = ~d1 & (h1 | c1);
        if ((int)(w0 | w1) < 0) { hitcount += j; oo2 ^= 1; }
      } else {
        float a0 = __uint_as_float((unsigned)A.x), a1 = __uint_as_float((unsigned)(A.x >> 32));
        float b0 = __uint_as_float((unsigned)A.y), b1 = __uint_as_float((unsigned)(A.y >> 32));
        float c0f = __uint_as_float((unsigned)B.x), c1f = __uint_as_float((unsigned)(B.x >> 32));
        float k0 = __uint_as_float((unsigned)B.y), k1 = __uint_as_float((unsigned)(B.y >> 32));
        float hb0 = fmaf(c0f, dz, fmaf(b0, dy, fmaf(a0, dx, od)));
        float hb1 = fmaf(c1f, dz, fmaf(b1, dy, fmaf(a1, dx, od)));
        float cc0 = fmaf(c0f, oz, fmaf(b0, oy, fmaf(a0, ox, oo))) + k0;
        float cc1 = fmaf(c1f, oz, fmaf(b1, oy, fmaf(a1, ox, oo))) + k1;
        float ds0 = fmaf(hb0, hb0, -cc0), ds1 = fmaf(hb1, hb1, -cc1);
        unsigned w0 = ~__float_as_uint(ds0) & (__float_as_uint(hb0) | __float_as_uint(cc0));
        unsigned w1 = ~__float_as_uint(ds1) & (__float_as_uint(hb1) | __float_as_uint(cc1));
        if ((int)(w0 | w1) < 0) { hitcount += j; oo += 1.f; }
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)hitcount + lo(oo2) + oo;
}
// ---- clock: SM cycles per ns under load
__global__ void k_clock(unsigned long long* res, int iters) {
  float a = threadIdx.x;
  unsigned long long t0, c0, t1, c1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  c0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 64; ++i) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a));
  c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { res[0] = c1 - c0; res[1] = t1 - t0; }
  if (a == 123.456f) res[2] = 1;
}

template <class F>
static float timeit(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s sm_%d%d SMs=%d clockRate=%d kHz smemOptin=%zu regsPerSM=%d l2=%d\n", pr.name, pr.major, pr.minor,
         pr.multiProcessorCount, clk_khz, pr.sharedMemPerBlockOptin, pr.regsPerMultiprocessor, pr.l2CacheSize);
  const int SM = pr.multiProcessorCount;
  float* out; CK(cudaMalloc(&out, sizeof(float) * SM * 8 * 1024));
  unsigned long long* res; CK(cudaMallocManaged(&res, 64));
  // clock under load
  k_clock<<<SM * 4, 256>>>(res, 200000); CK(cudaDeviceSynchronize());
  double ghz = (double)res[0] / (double)res[1];
  printf("clock under FFMA load: %.3f GHz (%llu cyc / %llu ns)\n", ghz, res[0], res[1]);

  const int iters = 20000;
  for (int wps : {4, 8, 16}) {       // warps per SM sub-partition set via blocks*threads
    int threads = 256, blocks = SM * (wps * 4 * 32 / threads);
    double winst = (double)blocks * (threads / 32) * iters * 32.0;  // warp instrs (32 per iter)
    float ms;
    ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    printf("FFMA   wps=%2d: %.3f ms  %.1f TFLOP/s  %.3f warp-inst/clk/SM @%.3fGHz\n", wps, ms,
           winst * 32 * 2 / ms * 1e-9, winst / (ms * 1e-3 * ghz * 1e9) / SM, ghz);
    ms = timeit([&] { k_ffma2<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    printf("FFMA2  wps=%2d: %.3f ms  %.1f TFLOP/s  %.3f warp-inst/clk/SM\n", wps, ms, winst * 32 * 4 / ms * 1e-9,
           winst / (ms * 1e-3 * ghz * 1e9) / SM);
    ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters / 4, 1.0001, 0.5); });
    printf("DFMA   wps=%2d: %.3f ms  %.1f TFLOP/s  %.3f warp-inst/clk/SM\n", wps, ms, winst / 4 * 32 * 2 / ms * 1e-9,
           winst / 4 / (ms * 1e-3 * ghz * 1e9) / SM);
    ms = timeit([&] { k_mufu<<<blocks, threads>>>(out, iters / 4); });
    printf("MUFU   wps=%2d: %.3f ms  %.3f warp-inst/clk/SM\n", wps, ms, winst / 4 / (ms * 1e-3 * ghz * 1e9) / SM);
  }
  // scan-shaped mix
  const int npairs = 244;
  size_t smem = npairs * 32;
  for (int wps : {2, 4, 8, 12, 16}) {
    int threads = 128, blocks = SM * (wps * 4 * 32 / threads);
    int mit = 200;
    double tests = (double)blocks * threads * mit * npairs * 2;
    float ms0 = timeit([&] { k_mix<0><<<blocks, threads, smem>>>(out, mit, npairs, 0.3f, 0.4f, 0.5f); });
    float ms1 = timeit([&] { k_mix<1><<<blocks, threads, smem>>>(out, mit, npairs, 0.3f, 0.4f, 0.5f); });
    double cyc0 = ms0 * 1e-3 * ghz * 1e9 * SM * 128 / tests, cyc1 = ms1 * 1e-3 * ghz * 1e9 * SM * 128 / tests;
    printf("MIX wps=%2d: packed %.3f ms = %.2f lane-cyc/test = %.1f TF(18/test) | scalar %.3f ms = %.2f lane-cyc/test = %.1f TF\n",
           wps, ms0, cyc0, tests * 18 / ms0 * 1e-9, ms1, cyc1, tests * 18 / ms1 * 1e-9);
  }
  k_clock<<<SM * 4, 256>>>(res, 200000); CK(cudaDeviceSynchronize());
  printf("clock under FFMA load (end): %.3f GHz\n", (double)res[0] / (double)res[1]);
  return 0;
}
