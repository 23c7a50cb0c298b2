// Internal helpers shared by the libsnmfnat translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <stdexcept>

#include "../../include/snmfnat.h"

namespace snmfnat {

// ---------------------------------------------------------------- errors
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

std::string vformat(const char* fmt, va_list ap);
[[noreturn]] void fail(int code, const char* fmt, ...);
void set_last_error(const char* msg);

#define SN_CUDA(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      ::snmfnat::fail(SNMFNAT_ECUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define SN_CUFFT(expr)                                                                          \
  do {                                                                                          \
    cufftResult _e = (expr);                                                                    \
    if (_e != CUFFT_SUCCESS)                                                                    \
      ::snmfnat::fail(SNMFNAT_ECUDA, "%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #expr, (int)_e); \
  } while (0)

#define SN_REQUIRE(cond, code, ...)               \
  do {                                            \
    if (!(cond)) ::snmfnat::fail(code, __VA_ARGS__); \
  } while (0)

// Wrap the body of an extern "C" entry point: no C++ exception may cross the C boundary.
#define SN_API_BEGIN try {
#define SN_API_END                                    \
  }                                                   \
  catch (const ::snmfnat::Error& e) {                 \
    ::snmfnat::set_last_error(e.what());              \
    return e.code;                                    \
  }                                                   \
  catch (const std::bad_alloc&) {                     \
    ::snmfnat::set_last_error("out of host memory");  \
    return SNMFNAT_ENOMEM;                            \
  }                                                   \
  catch (const std::exception& e) {                   \
    ::snmfnat::set_last_error(e.what());              \
    return SNMFNAT_EINVAL;                            \
  }                                                   \
  return SNMFNAT_OK;

// ---------------------------------------------------------------- device buffers
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    if (count == 0) return;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) fail(SNMFNAT_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    n = count;
  }
  void zero(cudaStream_t st) {
    if (n) SN_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
  }
};

// padded leading dimension of every F-long device vector / basis column (64-byte aligned columns)
inline int pad_ld(int F) { return (F + 7) / 8 * 8; }

}  // namespace snmfnat

// ---------------------------------------------------------------- context
struct snmfnat_ctx {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int max_smem_optin = 0;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  std::string last_error;
};

// ---------------------------------------------------------------- NVTX ranges (header-only NVTX3; no-ops without a tool)
#include <nvtx3/nvToolsExt.h>
namespace snmfnat {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace snmfnat

namespace snmfnat {
inline void count_launch(snmfnat_ctx* ctx, int n = 1) { ctx->launches += n; }
void check_launch(snmfnat_ctx* ctx, const char* what);
}  // namespace snmfnat

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
namespace snmfnat {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum in a fixed order (deterministic, identical on every CTA that runs the same code on the
// same data).  `scratch` holds >= 32 doubles.  Result is returned to every thread.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += scratch[i];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = scratch[0];
  for (int i = 1; i < nw; ++i) t = fmax(t, scratch[i]);
  return t;
}

}  // namespace snmfnat
#endif
