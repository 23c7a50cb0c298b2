// Micro-benchmarks on one SM: latency of a dependent mma.sync.m8n8k4.f64 chain, throughput with several chains / warps,
// and the issue cost of st.async to a peer CTA.   nvcc -arch=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_dmma(double* out, long long* clk, int iters) {
  double c[CH][2];
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// DFMA dependent chain
__global__ void k_dfma(double* out, long long* clk, int iters) {
  double x = threadIdx.x * 1e-9, y = 1.0000001;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) x = fma(x, y, 1e-9);
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

// st.async: every lane of `nw` warps sends `n` doubles to the peer CTA of a 2-CTA cluster; time until the peer's barrier completes
__global__ void __cluster_dims__(2, 1, 1) k_stas(long long* clk, int n, int vec) {
  __shared__ __align__(16) double buf[4096 + 1024];
  __shared__ unsigned long long bar;
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  const unsigned barA = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cl.sync();
  const unsigned total = blockDim.x * n * 8u * (vec ? 2u : 1u);
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barA), "r"(total) : "memory");
  cl.sync();
  const long long t0 = clock64();
  unsigned rb, ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(barA), "r"(rank ^ 1u));
  const unsigned la = (unsigned)__cvta_generic_to_shared(buf) + threadIdx.x * (vec ? 16u : 8u);
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank ^ 1u));
  for (int i = 0; i < n; ++i) {
    const unsigned dst = ra + (unsigned)i * blockDim.x * (vec ? 16u : 8u);
    if (vec)
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(dst), "l"(1ll), "l"(2ll), "r"(rb) : "memory");
    else
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(1ll), "r"(rb) : "memory");
  }
  const long long t1 = clock64();
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(barA), "r"(0u) : "memory");
  } while (!ok);
  const long long t2 = clock64();
  if (threadIdx.x == 0 && rank == 0) {
    clk[0] = t1 - t0;
    clk[1] = t2 - t0;
  }
  cl.sync();
}

int main() {
  double* out;
  long long* clk;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&clk, 4096);
  long long h[8];
  const int iters = 2000;
#define RUN_DMMA(CH, NW)                                                                                        \
  k_dmma<CH><<<1, 32 * NW>>>(out, clk, iters);                                                                  \
  cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);                                                                \
  printf("dmma chains/warp %d warps %d : %.1f clk per dmma per warp, SM rate %.2f clk per dmma\n", CH, NW,      \
         (double)h[0] / (iters * CH), (double)h[0] / (iters * CH * NW));
  RUN_DMMA(1, 1) RUN_DMMA(2, 1) RUN_DMMA(4, 1) RUN_DMMA(8, 1)
  RUN_DMMA(1, 4) RUN_DMMA(2, 4) RUN_DMMA(4, 4)
  RUN_DMMA(1, 8) RUN_DMMA(2, 8) RUN_DMMA(3, 8) RUN_DMMA(4, 8)
  RUN_DMMA(2, 16) RUN_DMMA(4, 16)
  k_dfma<<<1, 32>>>(out, clk, iters);
  cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);
  printf("dfma dependent chain: %.1f clk\n", (double)h[0] / iters);
  for (int vec = 0; vec < 2; ++vec)
    for (int nt : {32, 256})
      for (int n : {1, 4, 8}) {
        k_stas<<<2, nt>>>(clk, n, vec);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
        printf("st.async%s threads %d x %d stores (%d B): issue %lld clk, all landed %lld clk (%s)\n", vec ? ".v2" : "", nt, n,
               nt * n * (vec ? 16 : 8), h[0], h[1], cudaGetErrorString(e));
      }
  return 0;
}
