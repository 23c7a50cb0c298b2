// Micro-benchmark: measured FP64 peaks of this GPU (the roofline denominators of the online MU kernels):
// DFMA issue rate, DMMA (mma.sync m8n8k4 f64) rate, shared-memory read bandwidth (LDS.64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks_fp64 tools/peaks_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void lds_kernel(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    s0 += sm[(idx) & 16383];
    s1 += sm[(idx + 1024) & 16383];
    s2 += sm[(idx + 2048) & 16383];
    s3 += sm[(idx + 3072) & 16383];
    idx += 4096;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<sms * 4, 512>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  double flops = 2.0 * 8 * iters * (double)sms * 4 * 512;
  printf("{\"dfma_tflops\": %.2f, ", flops / ms / 1e9);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    dmma_kernel<<<sms * 4, 512>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  flops = 2.0 * 256 * 8 * iters * (double)sms * 4 * 16;
  printf("\"dmma_tflops\": %.2f, ", flops / ms / 1e9);
  cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    lds_kernel<<<sms, 1024, 131072>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  double bytes = 8.0 * 4 * iters * (double)sms * 1024;
  printf("\"lds64_tbs\": %.2f, \"sms\": %d, \"clock_mhz\": %d, \"name\": \"%s\"}\n", bytes / ms / 1e9, sms,
         p.clockRate / 1000, p.name);
  return 0;
}
