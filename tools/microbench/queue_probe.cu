// queue_probe.cu -- which tcgen05 instruction of the ISSUER delays the tcgen05.ld of the scan warp on the same
// SM sub-partition?  4 scan warps run the product's load + pack loop; one elected lane of warp 4 (sub-partition 0)
// issues one 128x128x32 TF32 tile (4 MMAs) every `period` cycles, paced by the clock (no barrier waits), with
//   mode 0: nothing     1: the MMAs only     2: MMAs + tcgen05.commit after every tile     3: commit only (no MMAs)
//   4: MMAs + commit, issued from warp 5 (sub-partition 1)
// Periods below the 296 cycles the four MMAs take saturate the tensor pipe: the issuer then stalls in dispatch.
// Output: cycles per 128-column slot for each scan warp.
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "tc_collect.cuh"

using namespace uecrt;
using namespace uecrt::tc;

__global__ void __launch_bounds__(256, 1) k_probe(int iters, int mode, int period, long long* out, uint32_t* sink, volatile int* flag) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x < 32) tmem_alloc(smem_addr(&tmem_slot), 512);
  if (threadIdx.x == 0) { mbar_init(smem_addr(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x >= 128) {
    const int issuer_warp = mode == 4 ? 5 : 4;
    if ((int)(threadIdx.x >> 5) == issuer_warp && mode != 0) {
      if (elect_one()) {
        long long next = clock64();
        int i = 0;
        while (*flag == 0) {
          if (mode != 3) issue_tile_ts(tb + (i % 3) * 128, tb + 384, smem_addr(smem), 128);
          if (mode >= 2) mma_commit(smem_addr(&bar));
          ++i;
          next += period;
          while (clock64() < next) {}
        }
      }
    }
  } else {
    const uint32_t lane_taddr = tb + ((threadIdx.x & 96u) << 16);
    uint32_t acc = 0;
    int va[32], vb[32];
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t t = lane_taddr + (it % 3) * 128;
      tmem_ld32(t, va);
      tmem_ld32(t + 32, vb);
      tmem_ld_wait();
      const uint32_t w0 = sign_word(va);
      tmem_ld32(t + 64, va);
      const uint32_t w1 = sign_word(vb);
      tmem_ld32(t + 96, vb);
      tmem_ld_wait();
      const uint32_t w2 = sign_word(va), w3 = sign_word(vb);
      acc += w0 + w1 + w2 + w3;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + (threadIdx.x >> 5)] = t1 - t0;
    sink[blockIdx.x * 128 + threadIdx.x] = acc;
    __syncwarp();
    if (threadIdx.x == 0) { __threadfence_block(); *flag = 1; }
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d_out; uint32_t* d_sink; int* d_flag;
  cudaMalloc(&d_out, 4 * sizeof(long long)); cudaMalloc(&d_sink, 128 * 4); cudaMalloc(&d_flag, 4);
  const int iters = 4000;
  const char* names[5] = {"nothing", "MMAs only", "MMAs + commit per tile", "commit only", "MMAs + commit from sub-partition 1"};
  for (int period : {700, 450, 320, 250, 100})
    for (int mode = 0; mode < 5; ++mode) {
      cudaMemset(d_out, 0, 4 * sizeof(long long)); cudaMemset(d_flag, 0, 4);
      k_probe<<<1, 256, 128 * 128>>>(iters, mode, period, d_out, d_sink, d_flag);      // one CTA: the flag is per CTA
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[4];
      cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
      printf("one tile per %d cycles, %-36s cycles per slot, scan warp on sub-partition 0..3: %.0f %.0f %.0f %.0f\n", period, names[mode],
             (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters);
    }
  return 0;
}
