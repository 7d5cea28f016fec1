// Micro-benchmark for the stacked-B formulation of conv_chain.cuh: cycles per tcgen05.mma (kind::f16, SS, M = 128,
// K = 16) as a function of N, of the B descriptor's atom stride (SBO 1024 = packed, 3072 = dy parts interleaved over
// dx) and of the A start row (0 = atom aligned, 7 = the dx = -1 tap of the row ring).  One CTA per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_stack_bench mma_stack_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ntire2022_esr_b200/csrc/tc_common.cuh"
using namespace esr;

__device__ __forceinline__ void mma_hi(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(a_hi), "r"(b_hi));
}

__global__ void __launch_bounds__(128, 1) k(int n, int iters, int sbo, int a_row, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tmem_s;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16((uint32_t)n);
      const uint32_t hi_a = 0x40004040u, hi_b = 0x40004000u | (uint32_t)(sbo >> 4);
      const uint32_t a0 = base + a_row * 128, b0 = base + 24 * 1024;   // A: 18 KB row slot, B: up to 72 KB
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t a_lo0 = 0x10000u | (((a0 + (i % 3) * 128u) & 0x3FFFFu) >> 4);
        const uint32_t b_lo0 = 0x10000u | (((b0 + (i % 3) * 1024u) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_hi(tm, a_lo0 + 2 * ks, hi_a, b_lo0 + 2 * ks, hi_b, idesc, 1u);
      }
      long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int iters = 300;
  for (int grid : {1, 148})
    for (int n : {32, 64, 96, 128, 144, 192})
      for (int sbo : {1024, 3072})
        for (int a_row : {0, 7}) {
          long long h[2] = {0, 0};
          for (int rep = 0; rep < 2; ++rep) {
            k<<<grid, 128, 210 * 1024>>>(n, iters, sbo, a_row, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("grid=%3d N=%3d SBO=%4d a_row=%d: issue %.1f, complete %.1f cycles/MMA  (math floor %.0f, smem model %.0f)\n", grid, n, sbo, a_row,
                 (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), n / 2.0, 32 + n / 4.0);
        }
  return 0;
}
