// Micro-benchmark: tcgen05.ld (TMEM -> registers) latency / throughput as a function of the number of reading
// warps, the vector width and concurrent tcgen05.mma traffic.  One CTA on one SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ntire2022_esr_b200/csrc/tc_common.cuh"
using namespace esr;

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr));
}

// warps 0..nw-1 read; warp 31 (if mma_n > 0) keeps the tensor pipe busy with N = mma_n MMAs
template <int X>
__global__ void __launch_bounds__(1024, 1) k(int nw, int nld, int iters, int mma_n, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_s, 512);
  if (threadIdx.x == 0) stop = 0;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tmem_s;
  if (warp < nw) {
    uint32_t acc = 0;
    const uint32_t trow = tm + ((uint32_t)((warp & 3) * 32) << 16);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t v[3][X];
      for (int s = 0; s < nld && s < 3; ++s) ld<X>(trow + (uint32_t)(((warp >> 2) * 3 + s) * X) % 448u, v[s]);
      tmem_ld_wait();
      for (int s = 0; s < nld && s < 3; ++s)
#pragma unroll
        for (int e = 0; e < X; ++e) acc ^= v[s][e];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; stop = 1; }
    sink[threadIdx.x] = acc;
  } else if (warp == 31 && mma_n > 0) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16((uint32_t)mma_n);
      int n = 0;
      while (!stop && n < 100000) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16_ss(tm + 256, umma_desc_sw128(base + ks * 32), umma_desc_sw128(base + 32768 + ks * 32), idesc, 1u);
        n += 4;
      }
      out[1] = n;
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  uint32_t* sink; cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 200;
  for (int mma_n : {0, 192})
    for (int x : {16, 32})
      for (int nw : {1, 4, 8, 16})
        for (int nld : {1, 3}) {
          long long h[2] = {0, 0};
          cudaMemset(d, 0, 16);
          for (int rep = 0; rep < 2; ++rep) {
            if (x == 16) k<16><<<1, 1024, 100 * 1024>>>(nw, nld, iters, mma_n, d, sink);
            else k<32><<<1, 1024, 100 * 1024>>>(nw, nld, iters, mma_n, d, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          const double cyc = (double)h[0] / iters;
          printf("mma_n=%3d x%-2d warps=%2d loads/round=%d: %.0f cycles/round  -> %.1f B/clk per SM\n", mma_n, x, nw, nld, cyc,
                 (double)nw * nld * 32 * x * 4 / cyc);
        }
  return 0;
}
