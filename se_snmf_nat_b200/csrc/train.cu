// Offline dictionary training on one rank's shard of frames: host side of the snmfnat_train_* C ABI.
//
// Reference: run_basis_train.m:80-91 calls sparse_nmf (src/sparse_nmf.m:71-292) with W and H both updated, KL
// divergence, on the power spectrogram of the training corpus.  Here every rank holds T_local frames of V (and the
// matching columns of H); an iteration is
//     hphase2_kernel  H-update of the local frames (+ cost of the state it started from, sum(H',2), tail row)
//     wphase2_kernel  G = (V ./ (W*H')) * H''   partial over the local frames
//     reduce          fixed-order sum of the per-CTA partials into  acc = [G (F x Kp) | sum(H',2) (Kp)]
//     ncclAllReduce   acc over ranks (the only collective: (F*Kp + Kp) floats)
//     wupdate_kernel  W-update + column normalisation (sparse_nmf.m:214-243), replicated on every rank
// State is fp32 in HBM; the tensor cores see tf32 operands (W rounded to nearest in its operand copy, R rounded in
// the epilogue, H truncated by the MMA) and accumulate in fp32 (TMEM).
#include <cuda.h>
#include <nccl.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include "common.cuh"
#include "train_kernels2.cuh"

using namespace snmfnat;
using namespace snmfnat::train;

#include <dlfcn.h>

// NCCL is bound at run time, not at link time: a host process that also imports torch must end up with ONE libnccl
// (torch's bundled copy when it is already loaded, the system library otherwise).
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
const NcclApi& nccl() {
  static NcclApi api;
  static bool loaded = false;
  if (!loaded) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) ::snmfnat::fail(SNMFNAT_ECUDA, "cannot load libnccl.so.2: %s", dlerror());
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy || !api.GetErrorString)
      ::snmfnat::fail(SNMFNAT_ECUDA, "libnccl.so.2 lacks a required symbol");
    loaded = true;
  }
  return api;
}
}  // namespace

#define SN_NCCL(expr)                                                                                          \
  do {                                                                                                         \
    ncclResult_t _e = (expr);                                                                                  \
    if (_e != ncclSuccess)                                                                                     \
      ::snmfnat::fail(SNMFNAT_ECUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl().GetErrorString(_e)); \
  } while (0)

struct snmfnat_train {
  snmfnat_ctx* ctx = nullptr;
  int F = 0, K = 0, Kp = 0, nkb = 0, ldv = 0;
  int64_t T = 0;
  double sparsity = 0.0;
  int tail_row = -1, nchunk = 0, ngroups = 0, grid_h = 0, ntiles = 0;
  int64_t ldt = 0;
  DevBuf<float> V, Vt, H[2], W0, Wm[2], Wt[2], invden[2], wtail[2], wn, acc, hs_part, gt_part, Gpart;
  DevBuf<double> cost_part, scal;
  bool probe = false;  // SNMFNAT_TRAIN_DEBUG=1: the kernels print the MMA issuer's wait/issue clocks of the first iteration
  int cur_h = 0, cur_w = 0;
  // TMA views of H and of the operand copy of W (train_kernels2.cuh): 128-row K-major tiles and 16-row MN-major slices
  CUtensorMap mH128[2], mW128[2], mHm16[2], mWm16[2];
  int nblk_h = 0, nlast_h = 0, nblocks_w = 0, nu = 2;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  double* h_scal = nullptr;  // pinned: [0] = div
  float* h_hs = nullptr;     // pinned: [Kp] sum(H,2) of the current H
  bool hs_valid = false;
  bool h_biased = false;  // H[cur_h] is in the biased storage form
  int64_t iters_done = 0;
};

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    SN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    SN_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, SNMFNAT_ECUDA,
               "cuTensorMapEncodeTiled is not available from this driver");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor [outer][inner] with `pitch` bytes between rows, box {32 floats, box_rows}, 128-byte swizzle,
// out-of-range elements read as zero / are not written.
void make_map(CUtensorMap* m, const float* base, uint64_t inner, uint64_t outer, uint64_t pitch, uint32_t box_rows,
              CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {(cuuint32_t)KB, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SN_REQUIRE(r == CUDA_SUCCESS, SNMFNAT_ECUDA, "cuTensorMapEncodeTiled failed with %d (inner %llu outer %llu pitch %llu)",
             (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch);
}

// 3-D view of a [outer][Kp] fp32 matrix as {32 floats, outer rows, Kp/32 column blocks}, box {32, box_rows, nkb}: one
// TMA instruction delivers box_rows rows of EVERY column block, stacked column block by column block.
void make_map_slices(CUtensorMap* m, const float* base, int nkb, uint64_t outer, uint64_t pitch, uint32_t box_rows,
                     CUtensorMapSwizzle swz) {
  cuuint64_t dims[3] = {(cuuint64_t)KB, outer, (cuuint64_t)nkb};
  cuuint64_t strides[2] = {pitch, (cuuint64_t)KB * 4};
  cuuint32_t box[3] = {(cuuint32_t)KB, box_rows, (cuuint32_t)nkb};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SN_REQUIRE(r == CUDA_SUCCESS, SNMFNAT_ECUDA, "cuTensorMapEncodeTiled (3-D slices) failed with %d (outer %llu pitch %llu nkb %d)",
             (int)r, (unsigned long long)outer, (unsigned long long)pitch, nkb);
}

__device__ __forceinline__ double blk_sum256(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}

__device__ __forceinline__ float tf32_rn_f(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// sparse_nmf.m:157-160: wn = sqrt(sum(w.^2)); w = w ./ wn  (h is rescaled by scale_h_kernel)
__global__ void init_w_kernel(const float* __restrict__ w0, int F, int K, int Kp, double sparsity, int tail_row,
                              float* __restrict__ Wm, float* __restrict__ Wt, float* __restrict__ invden,
                              float* __restrict__ wtail, float* __restrict__ wn) {
  __shared__ double sh[8];
  const int k = blockIdx.x;
  double ss = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const double w = w0[(size_t)k * F + f];
    ss += w * w;
  }
  const double nrm = sqrt(blk_sum256(ss, sh));
  double cs = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const float w = (float)((double)w0[(size_t)k * F + f] / nrm);
    Wm[(size_t)k * F + f] = w;
    Wt[(size_t)f * Kp + k] = tf32_rn_f(w);
    cs += w;
    if (f == tail_row) wtail[k] = w;
  }
  cs = blk_sum256(cs, sh);
  if (threadIdx.x == 0) {
    invden[k] = (float)(1.0 / fmax(cs + sparsity, 1e-9));  // sparse_nmf.m:192-193
    wn[k] = (float)nrm;
  }
}

// sparse_nmf.m:160 h = h .* wn', and conversion of the raw fp32 values to the biased storage form (train_kernels.cuh)
__global__ void scale_h_kernel(float* __restrict__ H, const float* __restrict__ wn, int K, int Kp, long long T,
                               int in_biased) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * Kp) return;
  const int k = (int)(i % Kp);
  float x = H[i];
  if (in_biased) x = h_unbias(x);
  x = (k < K) ? x * wn[k] : 0.f;
  H[i] = h_bias(x);
}

// acc = [G | hs]: fixed-order sums of the per-CTA partials of the two phase kernels
__global__ void reduce_kernel(const float* __restrict__ Gpart, int ngroups, int mma_rows_padded, int mma_rows,
                              const float* __restrict__ hs_part, const float* __restrict__ gt_part, int grid_h, int F,
                              int Kp, int tail_row, float* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = F * Kp + Kp;
  if (i >= total) return;
  const int f = i / Kp, k = i % Kp;
  float s = 0.f;
  if (f < mma_rows) {
    for (int g = 0; g < ngroups; ++g) s += Gpart[((size_t)g * mma_rows_padded + f) * Kp + k];
  } else if (f == tail_row) {
    for (int c = 0; c < grid_h; ++c) s += gt_part[(size_t)c * Kp + k];
  } else if (f == F) {
    for (int c = 0; c < grid_h; ++c) s += hs_part[(size_t)c * Kp + k];
  }
  acc[i] = s;
}

__global__ void cost_reduce_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = blk_sum256(s, sh);
  if (threadIdx.x == 0) out[0] = s;
}

// sparse_nmf.m:214-243 (KL): one block per atom
__global__ void wupdate_kernel(const float* __restrict__ acc, const float* __restrict__ Wc, int F, int K, int Kp,
                               double sparsity, int tail_row, float* __restrict__ Wn, float* __restrict__ Wt,
                               float* __restrict__ invden, float* __restrict__ wtail) {
  extern __shared__ float wnew[];
  __shared__ double sh[8];
  const int k = blockIdx.x;
  const double hs = acc[(size_t)F * Kp + k];
  double sa = 0.0, sw = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const double w = Wc[(size_t)k * F + f];
    sa += (double)acc[(size_t)f * Kp + k] * w;
    sw += w;
  }
  sa = blk_sum256(sa, sh);
  sw = blk_sum256(sw, sh);
  double nn = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const double w = Wc[(size_t)k * F + f];
    const double g = acc[(size_t)f * Kp + k];
    const double dpw = fmax(hs + w * sa, 1e-9);
    const double dmw = g + w * (hs * sw);
    const double x = w * dmw / dpw;
    wnew[f] = (float)x;
    nn += x * x;
  }
  nn = blk_sum256(nn, sh);
  const double inv = 1.0 / sqrt(nn);
  double cs = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const float w = (float)((double)wnew[f] * inv);
    Wn[(size_t)k * F + f] = w;
    Wt[(size_t)f * Kp + k] = tf32_rn_f(w);
    cs += w;
    if (f == tail_row) wtail[k] = w;
  }
  cs = blk_sum256(cs, sh);
  if (threadIdx.x == 0) invden[k] = (float)(1.0 / fmax(cs + sparsity, 1e-9));
}

// V [T][ldv] -> Vt [F][ldt]
__global__ void transpose_v_kernel(const float* __restrict__ V, int ldv, int F, long long T, float* __restrict__ Vt,
                                   long long ldt) {
  __shared__ float tile[32][33];
  const long long t0 = (long long)blockIdx.x * 32;
  const int f0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long t = t0 + i;
    const int f = f0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && f < F) ? V[t * ldv + f] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int f = f0 + i;
    const long long t = t0 + threadIdx.x;
    if (f < F && t < T) Vt[(long long)f * ldt + t] = tile[threadIdx.x][i];
  }
}

void launch_hphase(snmfnat_train* t, int update, int want_cost) {
  HPhase2Args a;
  a.F = t->F; a.Fm = t->tail_row >= 0 ? t->F - 1 : t->F; a.Kp = t->Kp; a.nkb = t->nkb;
  a.nh = t->nblk_h; a.nlast = t->nlast_h;
  a.ntiles = t->ntiles;
  a.update = update; a.want_cost = want_cost; a.tail_row = t->tail_row;
  a.T = t->T; a.ldt = t->ldt;
  a.Vt = t->Vt.p;
  a.invden = t->invden[t->cur_w].p; a.wtail = t->wtail[t->cur_w].p;
  a.hs_part = t->hs_part.p; a.gt_part = t->gt_part.p; a.cost_part = t->cost_part.p;
  a.probe = (update && t->iters_done == 0 && t->probe) ? 1 : 0;
  a.nu = t->nu;
  const int ch = t->cur_h, cw = t->cur_w;
  hphase2_kernel<<<t->grid_h, THREADS, phase2_smem_bytes(t->nkb, t->nu), t->ctx->stream>>>(t->mH128[ch], t->mH128[ch ^ 1],
                                                                                            t->mW128[cw], t->mWm16[cw], a);
  count_launch(t->ctx);
  check_launch(t->ctx, "hphase2_kernel");
}

void launch_wphase(snmfnat_train* t, int hbuf) {
  WPhase2Args a;
  a.F = t->F; a.Kp = t->Kp; a.nkb = t->nkb;
  a.nchunk = t->nchunk; a.ngroups = t->ngroups; a.nblocks = t->nblocks_w; a.ldv = t->ldv; a.T = t->T;
  a.V = t->V.p; a.Gpart = t->Gpart.p;
  a.nu = t->nu;
  a.probe = (t->iters_done == 0 && t->probe) ? 1 : 0;
  const int grid = t->nchunk * t->ngroups, cw = t->cur_w;
  wphase2_kernel<<<grid, THREADS, phase2_smem_bytes(t->nkb, t->nu), t->ctx->stream>>>(t->mW128[cw], t->mH128[hbuf],
                                                                                       t->mHm16[hbuf], a);
  count_launch(t->ctx);
  check_launch(t->ctx, "wphase2_kernel");
}

// div of the state the last hphase started from, summed over ranks, on the host (synchronises the stream)
double fetch_div(snmfnat_train* t) {
  cudaStream_t st = t->ctx->stream;
  cost_reduce_kernel<<<1, 256, 0, st>>>(t->cost_part.p, t->grid_h, t->scal.p);
  count_launch(t->ctx);
  if (t->world > 1) SN_NCCL(nccl().AllReduce(t->scal.p, t->scal.p, 1, ncclDouble, ncclSum, t->comm, st));
  SN_CUDA(cudaMemcpyAsync(t->h_scal, t->scal.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  return t->h_scal[0];
}

double current_hsum(snmfnat_train* t) {
  double s = 0.0;
  for (int k = 0; k < t->K; ++k) s += t->h_hs[k];
  return s;
}

// W-phase + reduction + all-reduce + W-update for the freshly written H buffer `hbuf`; swaps the buffers.
void finish_iteration(snmfnat_train* t) {
  cudaStream_t st = t->ctx->stream;
  const int hbuf = t->cur_h ^ 1;
  launch_wphase(t, hbuf);
  const int total = t->F * t->Kp + t->Kp;
  const int mma_rows = t->tail_row >= 0 ? t->F - 1 : t->F;
  reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(t->Gpart.p, t->ngroups, t->nchunk * BM, mma_rows, t->hs_part.p,
                                                      t->gt_part.p, t->grid_h, t->F, t->Kp, t->tail_row, t->acc.p);
  count_launch(t->ctx);
  if (t->world > 1) SN_NCCL(nccl().AllReduce(t->acc.p, t->acc.p, (size_t)total, ncclFloat, ncclSum, t->comm, st));
  const int nw = t->cur_w ^ 1;
  wupdate_kernel<<<t->K, 256, t->F * sizeof(float), st>>>(t->acc.p, t->Wm[t->cur_w].p, t->F, t->K, t->Kp, t->sparsity,
                                                           t->tail_row, t->Wm[nw].p, t->Wt[nw].p, t->invden[nw].p,
                                                           t->wtail[nw].p);
  count_launch(t->ctx);
  check_launch(t->ctx, "wupdate_kernel");
  SN_CUDA(cudaMemcpyAsync(t->h_hs, t->acc.p + (size_t)t->F * t->Kp, t->Kp * sizeof(float), cudaMemcpyDeviceToHost, st));
  t->hs_valid = true;
  t->cur_h = hbuf;
  t->cur_w = nw;
  t->iters_done++;
}

void do_reset(snmfnat_train* t) {
  cudaStream_t st = t->ctx->stream;
  init_w_kernel<<<t->K, 256, 0, st>>>(t->W0.p, t->F, t->K, t->Kp, t->sparsity, t->tail_row, t->Wm[0].p, t->Wt[0].p,
                                      t->invden[0].p, t->wtail[0].p, t->wn.p);
  count_launch(t->ctx);
  const long long n = (long long)t->T * t->Kp;
  scale_h_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t->H[t->cur_h].p, t->wn.p, t->K, t->Kp, t->T,
                                                                t->h_biased ? 1 : 0);
  t->h_biased = true;
  count_launch(t->ctx);
  check_launch(t->ctx, "train reset");
  t->cur_w = 0;
  t->hs_valid = false;
  t->iters_done = 0;
}

}  // namespace

extern "C" {

int snmfnat_train_create(snmfnat_ctx* ctx, int F, int K, int64_t T_local, double sparsity, int precision,
                         snmfnat_train** out) {
  SN_API_BEGIN
  SN_REQUIRE(ctx && out, SNMFNAT_EINVAL, "ctx/out is NULL");
  *out = nullptr;
  SN_REQUIRE(precision == 1, SNMFNAT_EUNSUPPORTED,
             "snmfnat_train_* is the tf32 tensor-core path (precision 1); use snmfnat_sparse_nmf for float64");
  SN_REQUIRE(F >= 1 && F <= 65535 && K >= 1 && T_local >= 1, SNMFNAT_EINVAL, "bad shape F=%d K=%d T=%lld", F, K,
             (long long)T_local);
  SN_REQUIRE(K <= MAX_KP, SNMFNAT_EUNSUPPORTED, "rank %d > %d is not supported by the tensor-memory layout", K, MAX_KP);
  SN_REQUIRE(T_local < (1LL << 31) - 256, SNMFNAT_EINVAL, "T_local too large for 32-bit TMA coordinates");
  SN_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<snmfnat_train> t(new snmfnat_train());
  t->ctx = ctx;
  t->F = F; t->K = K; t->T = T_local; t->sparsity = sparsity;
  t->Kp = (K + KB - 1) / KB * KB;
  t->nkb = t->Kp / KB;
  t->ldv = (F + 3) / 4 * 4;
  t->tail_row = (F > 1 && F % BM == 1) ? F - 1 : -1;
  const int mma_rows = t->tail_row >= 0 ? F - 1 : F;
  t->nchunk = (mma_rows + BM - 1) / BM;
  t->ntiles = (int)((T_local + BM - 1) / BM);
  t->ldt = (T_local + 3) / 4 * 4;
  t->grid_h = std::min(t->ntiles, ctx->sm_count);
  t->nu = phase2_units(t->nkb, (size_t)ctx->max_smem_optin);
  const size_t smem = phase2_smem_bytes(t->nkb, t->nu);
  SN_REQUIRE(smem <= (size_t)ctx->max_smem_optin, SNMFNAT_EUNSUPPORTED, "shared memory: need %zu bytes, device offers %d",
             smem, ctx->max_smem_optin);
  const int Fm = t->tail_row >= 0 ? F - 1 : F;
  t->nblk_h = (Fm + HB - 1) / HB;
  t->nlast_h = (Fm - (t->nblk_h - 1) * HB + 15) / 16 * 16;
  t->nblocks_w = (int)((T_local + HB - 1) / HB);
  t->ngroups = std::max(1, std::min(ctx->sm_count / t->nchunk, t->nblocks_w));
  SN_CUDA(cudaFuncSetAttribute(hphase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SN_CUDA(cudaFuncSetAttribute(wphase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = ctx->stream;
  t->V.alloc((size_t)T_local * t->ldv);
  t->Vt.alloc((size_t)F * t->ldt);
  for (int i = 0; i < 2; ++i) {
    t->H[i].alloc((size_t)T_local * t->Kp);
    t->H[i].zero(st);
    t->Wm[i].alloc((size_t)K * F);
    t->Wt[i].alloc((size_t)F * t->Kp);
    t->Wt[i].zero(st);
    t->invden[i].alloc(t->Kp);
    t->invden[i].zero(st);
    t->wtail[i].alloc(t->Kp);
    t->wtail[i].zero(st);
  }
  t->V.zero(st);
  t->W0.alloc((size_t)K * F);
  t->wn.alloc(t->Kp);
  t->acc.alloc((size_t)F * t->Kp + t->Kp);
  t->hs_part.alloc((size_t)t->grid_h * t->Kp);
  t->gt_part.alloc((size_t)t->grid_h * t->Kp);
  t->hs_part.zero(st);
  t->gt_part.zero(st);
  t->Gpart.alloc((size_t)t->ngroups * t->nchunk * BM * t->Kp);
  t->cost_part.alloc(t->grid_h);
  t->cost_part.zero(st);
  t->scal.alloc(2);
  t->probe = getenv("SNMFNAT_TRAIN_DEBUG") != nullptr;
  SN_CUDA(cudaMallocHost(&t->h_scal, 2 * sizeof(double)));
  SN_CUDA(cudaMallocHost(&t->h_hs, t->Kp * sizeof(float)));
  for (int i = 0; i < 2; ++i) {
    const uint64_t pitch = (uint64_t)t->Kp * 4;
    make_map(&t->mH128[i], t->H[i].p, t->Kp, T_local, pitch, BM);
    make_map(&t->mW128[i], t->Wt[i].p, t->Kp, F, pitch, BM);
    make_map_slices(&t->mHm16[i], t->H[i].p, t->nkb, T_local, pitch, SL, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    make_map_slices(&t->mWm16[i], t->Wt[i].p, t->nkb, F, pitch, SL, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  }
  SN_CUDA(cudaStreamSynchronize(st));
  *out = t.release();
  SN_API_END
}

int snmfnat_train_destroy(snmfnat_train* t) {
  SN_API_BEGIN
  if (t) {
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    if (t->comm) nccl().CommDestroy(t->comm);
    if (t->h_scal) cudaFreeHost(t->h_scal);
    if (t->h_hs) cudaFreeHost(t->h_hs);
    delete t;
  }
  SN_API_END
}

int snmfnat_train_nccl_unique_id(void* id128) {
  SN_API_BEGIN
  SN_REQUIRE(id128 != nullptr, SNMFNAT_EINVAL, "id128 is NULL");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  SN_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  SN_API_END
}

int snmfnat_train_attach_nccl(snmfnat_train* t, const void* nccl_unique_id, int rank, int world) {
  SN_API_BEGIN
  SN_REQUIRE(t && nccl_unique_id, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(world >= 1 && rank >= 0 && rank < world, SNMFNAT_EINVAL, "bad rank %d / world %d", rank, world);
  SN_CUDA(cudaSetDevice(t->ctx->device));
  if (t->comm) {
    nccl().CommDestroy(t->comm);
    t->comm = nullptr;
  }
  t->rank = rank;
  t->world = world;
  if (world > 1) {
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof(id));
    SN_NCCL(nccl().CommInitRank(&t->comm, world, id, rank));
  }
  SN_API_END
}

int snmfnat_train_get_layout(snmfnat_train* t, int* ldv, int* kp) {
  SN_API_BEGIN
  SN_REQUIRE(t != nullptr, SNMFNAT_EINVAL, "t is NULL");
  if (ldv) *ldv = t->ldv;
  if (kp) *kp = t->Kp;
  SN_API_END
}

// Rebuild the bin-major copy of V after the frame-major one ("V" of snmfnat_train_dev_ptr) was filled in place.
int snmfnat_train_commit_v(snmfnat_train* t) {
  SN_API_BEGIN
  SN_REQUIRE(t != nullptr, SNMFNAT_EINVAL, "t is NULL");
  SN_CUDA(cudaSetDevice(t->ctx->device));
  dim3 grid((unsigned)((t->T + 31) / 32), (unsigned)((t->F + 31) / 32)), block(32, 8);
  transpose_v_kernel<<<grid, block, 0, t->ctx->stream>>>(t->V.p, t->ldv, t->F, t->T, t->Vt.p, t->ldt);
  count_launch(t->ctx);
  check_launch(t->ctx, "transpose_v_kernel");
  SN_API_END
}

int snmfnat_train_set_data(snmfnat_train* t, const float* V, int v_on_device, const float* init_w, const float* init_h,
                           int h_on_device) {
  SN_API_BEGIN
  SN_REQUIRE(t != nullptr, SNMFNAT_EINVAL, "t is NULL");
  (void)v_on_device;
  (void)h_on_device;  // unified addressing: cudaMemcpyDefault resolves both
  SN_CUDA(cudaSetDevice(t->ctx->device));
  cudaStream_t st = t->ctx->stream;
  if (V)
    SN_CUDA(cudaMemcpy2DAsync(t->V.p, (size_t)t->ldv * 4, V, (size_t)t->F * 4, (size_t)t->F * 4, (size_t)t->T,
                              cudaMemcpyDefault, st));
  if (V) SN_REQUIRE(snmfnat_train_commit_v(t) == 0, SNMFNAT_ECUDA, "building the bin-major copy of V failed");
  if (init_h) {
    t->cur_h = 0;
    t->h_biased = false;
  }
  if (init_h)
    SN_CUDA(cudaMemcpy2DAsync(t->H[0].p, (size_t)t->Kp * 4, init_h, (size_t)t->K * 4, (size_t)t->K * 4, (size_t)t->T,
                              cudaMemcpyDefault, st));
  if (init_w) SN_CUDA(cudaMemcpyAsync(t->W0.p, init_w, (size_t)t->F * t->K * 4, cudaMemcpyDefault, st));
  SN_CUDA(cudaStreamSynchronize(st));
  if (init_w) do_reset(t);
  SN_API_END
}

void* snmfnat_train_dev_ptr(snmfnat_train* t, const char* which) {
  if (!t || !which) return nullptr;
  if (!std::strcmp(which, "V")) return t->V.p;                  // [T][ldv]
  if (!std::strcmp(which, "H")) {  // [T][Kp], to be FILLED with raw fp32 values before snmfnat_train_reset
    t->h_biased = false;
    return t->H[t->cur_h].p;
  }
  if (!std::strcmp(which, "W")) return t->Wm[t->cur_w].p;       // [K][F]
  if (!std::strcmp(which, "W_init")) return t->W0.p;            // [K][F] staging read by snmfnat_train_reset
  return nullptr;
}

int snmfnat_train_reset(snmfnat_train* t) {
  SN_API_BEGIN
  SN_REQUIRE(t != nullptr, SNMFNAT_EINVAL, "t is NULL");
  SN_CUDA(cudaSetDevice(t->ctx->device));
  do_reset(t);
  SN_API_END
}

int snmfnat_train_iterate(snmfnat_train* t, int n_iters, double* div, double* cost) {
  SN_API_BEGIN
  NvtxRange nvtx_it("snmfnat_train_iterate");
  SN_REQUIRE(t != nullptr && n_iters >= 0, SNMFNAT_EINVAL, "bad argument");
  SN_CUDA(cudaSetDevice(t->ctx->device));
  const bool want = div || cost;
  for (int i = 0; i < n_iters; ++i) {
    launch_hphase(t, 1, want && i > 0);
    if (want && i > 0) {
      const double d = fetch_div(t), c = d + t->sparsity * current_hsum(t);  // sparse_nmf.m:250,261
      if (div) div[i - 1] = d;
      if (cost) cost[i - 1] = c;
    }
    finish_iteration(t);
  }
  if (want && n_iters > 0) {
    launch_hphase(t, 0, 1);
    const double d = fetch_div(t), c = d + t->sparsity * current_hsum(t);
    if (div) div[n_iters - 1] = d;
    if (cost) cost[n_iters - 1] = c;
  }
  SN_API_END
}

int snmfnat_train_run(snmfnat_train* t, int max_iter, double conv_eps, double* div, double* cost, int* iters) {
  SN_API_BEGIN
  SN_REQUIRE(t != nullptr && max_iter >= 1, SNMFNAT_EINVAL, "bad argument");
  SN_CUDA(cudaSetDevice(t->ctx->device));
  // The cost of iteration `it` falls out of the first product of pass it+1; when the stop rule of
  // sparse_nmf.m:273-283 fires, that pass's H write-back is simply not adopted.
  double last = INFINITY;
  int done = 0;
  bool stopped = false;
  for (int it = 1; it <= max_iter + 1 && !stopped; ++it) {
    const bool last_pass = it == max_iter + 1;
    if (it > 1 || last_pass) {
      launch_hphase(t, last_pass ? 0 : 1, 1);
      const double d = fetch_div(t), c = d + t->sparsity * current_hsum(t);
      if (div) div[it - 2] = d;
      if (cost) cost[it - 2] = c;
      done = it - 1;
      if (it - 1 > 1 && conv_eps > 0 && std::fabs(c - last) / last < conv_eps) stopped = true;
      last = c;
      if (stopped || last_pass) break;
    } else {
      launch_hphase(t, 1, 0);
    }
    finish_iteration(t);
  }
  if (iters) *iters = done;
  SN_API_END
}

int snmfnat_train_get_w(snmfnat_train* t, float* w) {
  SN_API_BEGIN
  SN_REQUIRE(t && w, SNMFNAT_EINVAL, "NULL argument");
  SN_CUDA(cudaSetDevice(t->ctx->device));
  SN_CUDA(cudaMemcpyAsync(w, t->Wm[t->cur_w].p, (size_t)t->F * t->K * 4, cudaMemcpyDefault, t->ctx->stream));
  SN_CUDA(cudaStreamSynchronize(t->ctx->stream));
  SN_API_END
}

int snmfnat_train_get_acc(snmfnat_train* t, float* g, float* hs) {
  SN_API_BEGIN
  SN_REQUIRE(t != nullptr, SNMFNAT_EINVAL, "t is NULL");
  SN_REQUIRE(t->iters_done > 0, SNMFNAT_EINVAL, "no iteration has run yet");
  SN_CUDA(cudaSetDevice(t->ctx->device));
  cudaStream_t st = t->ctx->stream;
  // acc holds G as [F][Kp] (bin-major); the caller gets MATLAB's F x K column-major
  std::vector<float> tmp((size_t)t->F * t->Kp + t->Kp);
  SN_CUDA(cudaMemcpyAsync(tmp.data(), t->acc.p, tmp.size() * 4, cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  if (g)
    for (int k = 0; k < t->K; ++k)
      for (int f = 0; f < t->F; ++f) g[(size_t)k * t->F + f] = tmp[(size_t)f * t->Kp + k];
  if (hs)
    for (int k = 0; k < t->K; ++k) hs[k] = tmp[(size_t)t->F * t->Kp + k];
  SN_API_END
}

int snmfnat_train_get_h(snmfnat_train* t, float* h, int64_t t0, int64_t count) {
  SN_API_BEGIN
  SN_REQUIRE(t && h && t0 >= 0 && count >= 0 && t0 + count <= t->T, SNMFNAT_EINVAL, "bad range");
  SN_CUDA(cudaSetDevice(t->ctx->device));
  if (count)
    SN_CUDA(cudaMemcpy2DAsync(h, (size_t)t->K * 4, t->H[t->cur_h].p + (size_t)t0 * t->Kp, (size_t)t->Kp * 4,
                              (size_t)t->K * 4, (size_t)count, cudaMemcpyDefault, t->ctx->stream));
  SN_CUDA(cudaStreamSynchronize(t->ctx->stream));
  if (t->h_biased) {
    cudaPointerAttributes at;
    SN_REQUIRE(cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type != cudaMemoryTypeDevice, SNMFNAT_EINVAL,
               "snmfnat_train_get_h needs a host buffer");
    for (int64_t i = 0; i < count * t->K; ++i) h[i] = h_unbias(h[i]);
  }
  SN_API_END
}

}  // extern "C"
