// Which shared-memory layouts does tcgen05.mma kind::tf32 accept for K-major and MN-major operands?
// One CTA, TMA-loaded tiles, a table of descriptor variants; prints the relative error of each against the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu && tools/umma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../se_snmf_nat_b200/csrc/umma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Variant {
  int b_mn;         // 0: B K-major (D = A[128x64] * B[32x64]'), 1: B MN-major (D = A[128x32] * B[32x64])
  int layout;       // descriptor layout type for B (2 = SW128, 1 = SW128 base 32B)
  int lbo, sbo;     // bytes
  int kstep;        // bytes added to B's start address per k-step
  int a_tmem;       // A operand from TMEM
  int bmap;         // which tensor map loads B: 0 = SW128, 1 = SW128_ATOM_32B
};

__device__ __forceinline__ uint64_t desc_l(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB0,
             const __grid_constant__ CUtensorMap mapB1, Variant v, float* out) {
  using namespace umma;
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = smem;             // 2 column blocks x [128 x 128 B]
  uint8_t* Bs = As + 2 * 16384;   // 2 column blocks x [32 x 128 B]
  uint64_t* bars = (uint64_t*)(Bs + 2 * 4096);
  uint32_t* slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bars + 0, 2 * 16384 + 2 * 4096);
    for (int kb = 0; kb < 2; ++kb) tma_load_2d(As + kb * 16384, &mapA, bars + 0, kb * 32, 0);
    for (int kb = 0; kb < 2; ++kb) tma_load_2d(Bs + kb * 4096, v.bmap ? &mapB1 : &mapB0, bars + 0, kb * 32, 0);
  }
  mbar_wait(bars + 0, 0);
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
  if (v.a_tmem) {  // A' = first column block of A (128 x 32) -> TMEM columns 128..159
    uint32_t r[32];
    const int row = threadIdx.x;
    for (int j = 0; j < 32; ++j) r[j] = *(const uint32_t*)(As + sw128_off(row, j));
    tmem_st32(lane_addr + 128, r);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(As), b_addr = smem_u32(Bs);
    if (!v.b_mn) {
      const uint32_t id = idesc_tf32(128, 32, 0, 0);
      for (int k = 0; k < 8; ++k) {
        const uint64_t da = desc_l(a_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, 2);
        const uint64_t db = desc_l(b_addr + (k >> 2) * 4096 + (k & 3) * v.kstep, v.lbo, v.sbo, v.layout);
        mma_ss(tmem, da, db, id, k > 0);
      }
    } else {
      const uint32_t id = idesc_tf32(128, 64, 0, 1);
      for (int k = 0; k < 4; ++k) {
        const uint64_t db = desc_l(b_addr + k * v.kstep, v.lbo, v.sbo, v.layout);
        if (v.a_tmem) {
          mma_ts(tmem, tmem + 128 + 8 * k, db, id, k > 0);
        } else {
          const uint64_t da = desc_l(a_addr + k * 32, 16, 1024, 2);
          mma_ss(tmem, da, db, id, k > 0);
        }
      }
    }
    mma_commit(bars + 1);
  }
  mbar_wait(bars + 1, 0);
  tc_fence_after();
  uint32_t d[32];
  for (int cb = 0; cb < 2; ++cb) {
    tmem_ld32(lane_addr + cb * 32, d);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 64 + cb * 32 + j] = __uint_as_float(d[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)p;
  std::vector<float> A(128 * 64), B(32 * 64);
  srand(1);
  for (auto& x : A) x = (float)(rand() % 17 - 8) / 8.f;   // exactly representable in tf32
  for (auto& x : B) x = (float)(rand() % 13 - 6) / 4.f;
  float *dA, *dB, *dO;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dO, 128 * 64 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  auto mk = [&](CUtensorMap* m, float* base, int rows, CUtensorMapSwizzle sw) {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t st[1] = {64 * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  CUtensorMap mA, mB0, mB1;
  mk(&mA, dA, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  mk(&mB0, dB, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  mk(&mB1, dB, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  // host references
  std::vector<double> Rk(128 * 32), Rm(128 * 64);
  for (int i = 0; i < 128; ++i)
    for (int n = 0; n < 32; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)A[i * 64 + k] * B[n * 64 + k];
      Rk[i * 32 + n] = s;
    }
  for (int i = 0; i < 128; ++i)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int k = 0; k < 32; ++k) s += (double)A[i * 64 + k] * B[k * 64 + n];
      Rm[i * 64 + n] = s;
    }
  struct Named { const char* name; Variant v; };
  std::vector<Named> vs = {
      {"K-major  B, TMA SW128,      layout 2, sbo 1024, kstep 32", {0, 2, 16, 1024, 32, 0, 0}},
      {"K-major  B, TMA SW128_A32B, layout 1, sbo 1024, kstep 32", {0, 1, 16, 1024, 32, 0, 1}},
      {"K-major  B, TMA SW128_A32B, layout 1, sbo  512, kstep 32", {0, 1, 16, 512, 32, 0, 1}},
      {"K-major  B, TMA SW128_A32B, layout 2, sbo 1024, kstep 32", {0, 2, 16, 1024, 32, 0, 1}},
      {"MN-major B, TMA SW128,      layout 2, lbo 4096 sbo 1024, kstep 1024", {1, 2, 4096, 1024, 1024, 0, 0}},
      {"MN-major B, TMA SW128,      layout 2, lbo 1024 sbo 4096, kstep 1024", {1, 2, 1024, 4096, 1024, 0, 0}},
      {"MN-major B, TMA SW128_A32B, layout 1, lbo 4096 sbo  512, kstep 1024", {1, 1, 4096, 512, 1024, 0, 1}},
      {"MN-major B, TMA SW128_A32B, layout 1, lbo 4096 sbo 1024, kstep 1024", {1, 1, 4096, 1024, 1024, 0, 1}},
      {"MN-major B, TMA SW128_A32B, layout 1, lbo  512 sbo 4096, kstep 1024", {1, 1, 512, 4096, 1024, 0, 1}},
      {"MN-major B, TMA SW128_A32B, layout 1, lbo 1024 sbo 4096, kstep 1024", {1, 1, 1024, 4096, 1024, 0, 1}},
      {"MN-major B, TMA SW128_A32B, layout 1, lbo 4096 sbo  512, kstep 1024, A in TMEM", {1, 1, 4096, 512, 1024, 1, 1}},
      {"MN-major B, TMA SW128,      layout 2, lbo 4096 sbo 1024, kstep 1024, A in TMEM", {1, 2, 4096, 1024, 1024, 1, 0}},
  };
  std::vector<float> O(128 * 64);
  if (only >= (int)vs.size()) return 2;
  int vi = -1;
  for (auto& nv : vs) {
    if (++vi != only && only >= 0) continue;
    CK(cudaMemset(dO, 0, O.size() * 4));
    probe_kernel<<<1, 128, 48 * 1024>>>(mA, mB0, mB1, nv.v, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-90s -> CUDA error %s\n", nv.name, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    const int N = nv.v.b_mn ? 64 : 32;
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < N; ++n) {
        const double r = nv.v.b_mn ? Rm[i * 64 + n] : Rk[i * 32 + n];
        const double d = O[i * 64 + n] - r;
        num += d * d;
        den += r * r;
      }
    printf("%-90s -> rel err %.3e %s\n", nv.name, sqrt(num / den), sqrt(num / den) < 1e-5 ? "OK" : "");
  }
  return 0;
}
