// scan_probe.cu -- what bounds the SCAN role of the tensor-core scan (tc_collect.cuh scanner_role)?
// Times, per 128-column slot and per warp, on otherwise idle SMs:
//   L   the TMEM loads alone (4 x tcgen05.ld.32x32b.x32 + wait)
//   P   the sign packing alone (4 x sign_word on registers)
//   LP  the product's load/pack sequence
// with 1 and 2 warps per SM sub-partition (128 / 256 threads), and with MMAs running concurrently or not.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I uecraytracing_b200/csrc tools/microbench/scan_probe.cu -o build/scan_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "tc_collect.cuh"

using namespace uecrt;
using namespace uecrt::tc;

template <int MODE>
__global__ void __launch_bounds__(384, 1) k_probe(int iters, int with_mma, long long* out, uint32_t* sink) {
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
  const int nscan = blockDim.x - 128;          // the last warpgroup issues MMAs (or idles)
  if ((int)threadIdx.x >= nscan) {
    if (with_mma && (threadIdx.x & 127) < 32) {
      if (elect_one()) {
        // keep the tensor pipe busy writing TMEM columns [0, 384) for about as long as the scan loop runs
        for (int i = 0; i < iters * with_mma; ++i) {
          issue_tile_ts(tb + (i % 3) * 128, tb + 384, smem_addr(smem), 128);
          if ((i & 7) == 7) { mma_commit(smem_addr(&bar)); mbar_wait(smem_addr(&bar), (i >> 3) & 1); }
        }
      }
    }
  } else {
    const uint32_t lane_taddr = tb + ((threadIdx.x & 96u) << 16);
    uint32_t acc = 0;
    int va[32], vb[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { va[i] = threadIdx.x * 31 + i; vb[i] = threadIdx.x * 17 - i; }
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t t = lane_taddr + (it % 3) * 128;
      if (MODE == 0) {             // loads only
        tmem_ld32(t, va); tmem_ld32(t + 32, vb); tmem_ld_wait();
        acc += va[0] ^ vb[5];
        tmem_ld32(t + 64, va); tmem_ld32(t + 96, vb); tmem_ld_wait();
        acc += va[3] ^ vb[9];
      } else if (MODE == 1) {      // pack only
        const uint32_t w0 = sign_word(va), w1 = sign_word(vb);
        va[it & 31] += w0; vb[(it + 7) & 31] ^= w1;
        const uint32_t w2 = sign_word(va), w3 = sign_word(vb);
        va[(it + 3) & 31] ^= w2; vb[(it + 11) & 31] += w3;
        acc += w0 + w1 + w2 + w3;
      } else {                     // the product's sequence
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
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 148 * 8 * sizeof(long long)); cudaMalloc(&d_sink, 148 * 384 * 4);
  const int iters = 4000;
  const char* names[3] = {"loads only", "pack only", "product load+pack"};
  for (int mode = 0; mode < 3; ++mode)
    for (int wgs = 1; wgs <= 2; ++wgs)
      for (int mma = 0; mma <= 2; ++mma) {
        cudaMemset(d_out, 0, 148 * 8 * sizeof(long long));
        const int threads = wgs * 128 + 128;
        const size_t smem = 128 * 128;
        if (mode == 0) k_probe<0><<<148, threads, smem>>>(iters, mma, d_out, d_sink);
        if (mode == 1) k_probe<1><<<148, threads, smem>>>(iters, mma, d_out, d_sink);
        if (mode == 2) k_probe<2><<<148, threads, smem>>>(iters, mma, d_out, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148 * 8];
        cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        double sum = 0; int cnt = 0;
        for (int i = 0; i < 148 * 8; ++i) if (h[i]) { sum += (double)h[i]; ++cnt; }
        printf("%-18s scan warps/SMSP %d, concurrent MMAs per slot %d: %.0f cycles per 128-column slot per warp -> %.0f cycles per slot per SMSP\n",
               names[mode], wgs, mma * 1, sum / cnt / iters, sum / cnt / iters / wgs);
      }
  return 0;
}
