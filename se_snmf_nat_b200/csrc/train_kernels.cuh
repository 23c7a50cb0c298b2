// Definitions shared by the dictionary-training kernels (sparse_nmf with W and H both updated, KL divergence,
// src/sparse_nmf.m:186-286 as called from run_basis_train.m:84-88): tile constants, the tf32 rounding bias of H, small
// helpers.  The two warp-specialised tcgen05 kernels of a multiplicative-update iteration are in train_kernels2.cuh (the
// first-generation kernels that streamed 16-row chunks were removed in round 2: 128 against 326 TFLOP/s at 1.25 M frames,
// K = 256), the small element-wise kernels in train.cu.
#pragma once
#include <cstring>
#include "umma.cuh"

namespace snmfnat {
namespace train {

constexpr int BM = 128;      // rows of every MMA (TMEM lanes): frames (hphase) or bins (wphase)
constexpr int KB = 32;       // floats per swizzled column block (128 bytes)
constexpr int MAX_KP = 256;  // padded rank limit (TMEM: Kp accumulator columns + 2 Lambda buffers)
constexpr int TMEM_COLS = 512;
constexpr int LAM_COL = 256;
constexpr int THREADS = 32 * 10;  // warp 0: TMA, warp 1: MMA issue + TMEM alloc, warps 2-9: two epilogue groups
constexpr float FLRF = 1e-9f;

// H lives in HBM with half a tf32 ulp added to its bit pattern: the tensor core truncates fp32 operands to tf32, so
// it then sees round-to-nearest(h) instead of a value biased towards zero, and the exact fp32 state is recovered by
// subtracting the same constant (h >= 0 always; integer add / subtract is exactly reversible).
constexpr uint32_t H_BIAS = 0x1000u;
__host__ __device__ __forceinline__ float h_unbias(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(x) - H_BIAS);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  u -= H_BIAS;
  memcpy(&x, &u, 4);
  return x;
#endif
}
__device__ __forceinline__ float h_bias(float x) { return __uint_as_float(__float_as_uint(x) + H_BIAS); }

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return (uint8_t*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023);
}

}  // namespace train
}  // namespace snmfnat
