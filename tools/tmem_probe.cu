// Micro-benchmark: read throughput of tensor memory through tcgen05.ld (32x32b.x16 / .x32) from 1, 4 and 8 warps of one CTA,
// and shared-memory ld.shared.v2.f64 throughput beside it.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I se_snmf_nat_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#include "umma.cuh"
using namespace umma;

__global__ void k_tmem(long long* clk, unsigned* sink, int iters, int nwarps) {
  __shared__ uint32_t taddr_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&taddr_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = taddr_s + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t r[32];
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  for (int c = 0; c < 512; c += 32) tmem_st32(base + c, r);
  tmem_wait_st();
  __syncthreads();
  unsigned acc = 0;
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 512; c += 64) {
        uint32_t a[32], b[32];
        tmem_ld32(base + c, a);
        tmem_ld32(base + c + 32, b);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += a[i] ^ b[i];
      }
    }
  }
  const long long t1 = clock64();
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  if (warp == 0) tmem_dealloc(taddr_s, 512);
}

__global__ void k_smem(long long* clk, double* sink, int iters, int nwarps) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc = 0.0;
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const double2 v = *reinterpret_cast<const double2*>(sm + (size_t)c * 512 + 2 * lane + 64 * (warp & 3));
        acc += v.x + v.y;
      }
    }
  }
  const long long t1 = clock64();
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  long long* clk; unsigned* sink; double* dsink;
  cudaMalloc(&clk, 64); cudaMalloc(&sink, 4096); cudaMalloc(&dsink, 8192);
  const int iters = 200;
  for (int nw : {1, 2, 4, 8}) {
    k_tmem<<<1, 256>>>(clk, sink, iters, nw);
    long long h = 0;
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)nw * iters * 512 * 32 * 4;
    printf("tcgen05.ld 32x32b.x32, %d warp(s): %.1f B/clk per SM (%.1f per warp)  [%s]\n", nw, bytes / h, bytes / h / nw, cudaGetErrorString(e));
  }
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int nw : {1, 2, 4, 8}) {
    k_smem<<<1, 256, 65536>>>(clk, dsink, iters * 8, nw);
    long long h = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)nw * iters * 8 * 16 * 512;
    printf("ld.shared.v2.f64,      %d warp(s): %.1f B/clk per SM\n", nw, bytes / h);
  }
  // both at once would need one kernel; the two pipes are measured separately here
  return 0;
}
