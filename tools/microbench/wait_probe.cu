// wait_probe.cu -- does a WAITING warp slow down the scan warp that shares its SM sub-partition?
// 4 scan warps (one per sub-partition) run the product's TMEM load + sign-pack loop; one extra warp on
// sub-partition 0 waits on an mbarrier that completes only when the scan warps are done, in different ways:
//   0 no extra warp            1 mbar_wait (try_wait loop), all 32 lanes     2 mbar_wait, one elected lane
//   3 test_wait + nanosleep(100), one lane     4 one lane spinning on test_wait     5 one lane in a long nanosleep loop (no barrier)
// Output: cycles per 128-column slot for each scan warp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I uecraytracing_b200/csrc tools/microbench/wait_probe.cu -o build/wait_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "tc_collect.cuh"

using namespace uecrt;
using namespace uecrt::tc;

__global__ void __launch_bounds__(256, 1) k_probe(int iters, int mode, long long* out, uint32_t* sink, volatile int* flag) {
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x < 32) tmem_alloc(smem_addr(&tmem_slot), 512);
  if (threadIdx.x == 0) { mbar_init(smem_addr(&bar), 128); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x >= 128) {
    const bool my_warp = (threadIdx.x >> 5) == 4;        // warp 4 -> sub-partition 0
    if (my_warp && mode != 0) {
      if (mode == 1) mbar_wait(smem_addr(&bar), 0);
      else if (elect_one()) {
        if (mode == 2) mbar_wait(smem_addr(&bar), 0);
        if (mode == 3) while (!mbar_test(smem_addr(&bar), 0)) __nanosleep(100);
        if (mode == 4) while (!mbar_test(smem_addr(&bar), 0)) {}
        if (mode == 5) while (*flag == 0) __nanosleep(1000);
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
    mbar_arrive(smem_addr(&bar));
    if (threadIdx.x == 0) *flag = 1;
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d_out; uint32_t* d_sink; int* d_flag;
  cudaMalloc(&d_out, 148 * 4 * sizeof(long long)); cudaMalloc(&d_sink, 148 * 128 * 4); cudaMalloc(&d_flag, 4);
  const int iters = 4000;
  const char* names[6] = {"no extra warp", "mbar_wait, 32 lanes", "mbar_wait, 1 lane", "test_wait + nanosleep(100), 1 lane", "test_wait spin, 1 lane", "nanosleep(1000) loop on a global flag, 1 lane"};
  for (int mode = 0; mode < 6; ++mode) {
    cudaMemset(d_out, 0, 148 * 4 * sizeof(long long)); cudaMemset(d_flag, 0, 4);
    k_probe<<<148, 256>>>(iters, mode, d_out, d_sink, d_flag);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148 * 4];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    double w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 148 * 4; ++i) w[i & 3] += (double)h[i] / 148 / iters;
    printf("%-46s cycles per slot, scan warp on sub-partition 0..3: %.0f %.0f %.0f %.0f\n", names[mode], w[0], w[1], w[2], w[3]);
  }
  return 0;
}
