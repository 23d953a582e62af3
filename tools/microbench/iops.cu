// Integer-pipe throughput of the instructions the tensor-scan's bit packing can be built from (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o iops iops.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int OP>
__global__ void k_op(unsigned* out, int iters, unsigned x, unsigned y) {
  unsigned a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 2654435761u + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 0) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(a[i]) : "r"(x));            // SHF.L.W (funnel)
        if (OP == 1) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a[i]) : "r"(x));             // PRMT
        if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(x), "r"(y));   // LOP3
        if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));                      // IADD
        if (OP == 4) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));                      // VIMNMX
        if (OP == 5) asm volatile("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));     // IDP4A
        if (OP == 6) asm volatile("{.reg .pred p; setp.lt.s32 p, %1, 0; selp.u32 %0, %0, %2, p;}" : "+r"(a[i]) : "r"(x), "r"(y));  // ISETP + SEL
        if (OP == 7) asm volatile("{add.cc.u32 %1, %1, %1; addc.u32 %0, %0, %0;}" : "+r"(a[i]), "+r"(x));   // carry chain (2 instr)
        if (OP == 8) asm volatile("shl.b32 %0, %0, 1;" : "+r"(a[i]));                                // SHL
        if (OP == 9) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));       // IMAD
      }
  }
  unsigned s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
}
template <int OP> static void run(const char* name, unsigned* out, int SM) {
  const int iters = 20000, threads = 256, blocks = SM * 4;     // 8 warps per sub-partition
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_op<OP><<<blocks, threads>>>(out, iters, 0x80000001u, 0x01020408u); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); k_op<OP><<<blocks, threads>>>(out, iters, 0x80000001u, 0x01020408u); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double winst = (double)blocks * (threads / 32) * iters * 32.0;
  printf("%-14s %.3f ms  %.2f warp-inst/clk/SM (per asm statement)\n", name, ms, winst / (ms * 1e-3 * 1.965e9) / SM);
}
int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  unsigned* out; CK(cudaMalloc(&out, 4 * pr.multiProcessorCount * 4 * 256));
  const int SM = pr.multiProcessorCount;
  run<0>("SHF.L.W", out, SM); run<1>("PRMT", out, SM); run<2>("LOP3", out, SM); run<3>("IADD", out, SM); run<4>("VIMNMX", out, SM);
  run<5>("IDP4A", out, SM); run<6>("ISETP+SEL", out, SM); run<7>("ADD.CC+ADDC", out, SM); run<8>("SHL", out, SM); run<9>("IMAD", out, SM);
  return 0;
}
