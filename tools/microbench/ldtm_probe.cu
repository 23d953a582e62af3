// ldtm_probe.cu -- does a tcgen05.ld wait behind tcgen05.mma issued from the SAME SM sub-partition?
// 4 probe warps (one per sub-partition) time single tcgen05.ld.32x32b.x32 + wait::ld round trips with clock64
// while one elected lane of warp 4 + s (sub-partition s) issues 128x128x32 TF32 tiles (4 MMAs + commit) back to
// back, keeping `depth` tiles in flight (an mbarrier per tile bounds the backlog).  depth 0: no MMAs.
// Output: mean / max round trip per probe warp.  nvcc -arch=sm_100a -I uecraytracing_b200/csrc -o build/ldtm_probe tools/microbench/ldtm_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "tc_collect.cuh"

using namespace uecrt;
using namespace uecrt::tc;

__global__ void __launch_bounds__(256, 1) k_probe(int iters, int issuer_sp, int depth, long long* out, uint32_t* sink, volatile int* flag) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar[4];
  if (threadIdx.x < 32) tmem_alloc(smem_addr(&tmem_slot), 512);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_addr(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x >= 128) {
    if ((int)(threadIdx.x >> 5) == 4 + issuer_sp && depth > 0) {
      if (elect_one()) {
        uint32_t i = 0;
        while (*flag == 0) {
          if (i >= (uint32_t)depth) mbar_wait(smem_addr(&bar[(i - depth) & 3u]), ((i - depth) >> 2) & 1u);   // tile i - depth has completed
          issue_tile_ts(tb + (i % 3) * 128, tb + 384, smem_addr(smem), 128);
          mma_commit(smem_addr(&bar[i & 3u]));
          ++i;
        }
      }
    }
  } else {
    const uint32_t lane_taddr = tb + ((threadIdx.x & 96u) << 16);
    uint32_t acc = 0;
    int va[32];
    long long sum = 0, mx = 0;
    __syncwarp();
    for (int it = 0; it < iters; ++it) {
      const uint32_t t = lane_taddr + (it % 3) * 128;
      const long long t0 = clock64();
      tmem_ld32(t, va);
      tmem_ld_wait();
      const long long t1 = clock64();
      acc += (uint32_t)va[it & 31];
      sum += t1 - t0; mx = t1 - t0 > mx ? t1 - t0 : mx;
      for (int k = 0; k < 8; ++k) acc = acc * 1664525u + 1013904223u;     // some ALU work between loads
    }
    if ((threadIdx.x & 31) == 0) { out[(threadIdx.x >> 5) * 2] = sum; out[(threadIdx.x >> 5) * 2 + 1] = mx; }
    sink[threadIdx.x] = acc;
  }
  if (threadIdx.x < 128) { if (threadIdx.x == 0) { __threadfence_block(); *flag = 1; } }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d_out; uint32_t* d_sink; int* d_flag;
  cudaMalloc(&d_out, 8 * sizeof(long long)); cudaMalloc(&d_sink, 128 * 4); cudaMalloc(&d_flag, 4);
  const int iters = 20000;
  for (int sp : {0, 1})
    for (int depth : {0, 1, 2, 3}) {
      cudaMemset(d_out, 0, 8 * sizeof(long long)); cudaMemset(d_flag, 0, 4);
      k_probe<<<1, 256, 128 * 128>>>(iters, sp, depth, d_out, d_sink, d_flag);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[8];
      cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
      printf("issuer on sub-partition %d, %d tile(s) of 4 MMAs in flight: tcgen05.ld x32 round trip, probe warp on sub-partition 0..3: mean %.0f %.0f %.0f %.0f  max %lld %lld %lld %lld cycles\n",
             sp, depth, (double)h[0] / iters, (double)h[2] / iters, (double)h[4] / iters, (double)h[6] / iters, h[1], h[3], h[5], h[7]);
    }
  return 0;
}
