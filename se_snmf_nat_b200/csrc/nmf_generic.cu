// Generic float64 sparse_nmf on the GPU: any F, n, r, any beta-divergence, selective W / H updates, per-entry
// sparsity, and the missing-data-imputation variants.  Restates src/sparse_nmf.m:157-286, src/snmf_mdi.m:175,251-255,
// 297-303 and src/DNMF_adapt.m:3-20 with tiled FP64 GEMM kernels + fused elementwise kernels; the iteration loop is
// driven from the host (one scalar read back per iteration for the early-stop test).  This is the shape-agnostic path
// behind the L1 entry points; the online solves and at-scale training have their own specialised kernels.
#include <cmath>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace snmfnat {

// ------------------------------------------------------------------------------------------------- GEMM
// C[M x N] (column-major, ldc) = A'[M x K] * B'[K x N] where A'(i,k) = A[i*sai + k*sak], B'(k,j) = B[k*sbk + j*sbj].
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256)
gemm_f64_kernel(int M, int N, int K, const double* __restrict__ A, long sai, long sak, const double* __restrict__ B,
                long sbk, long sbj, double* __restrict__ C, int ldc) {
  __shared__ double As[GK][GT + 1];
  __shared__ double Bs[GK][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.x * GT, j0 = blockIdx.y * GT;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int t = threadIdx.x; t < GK * GT; t += 256) {
      const int kk = t / GT, ii = t % GT;
      const int gi = i0 + ii, gk = k0 + kk;
      As[kk][ii] = (gi < M && gk < K) ? A[gi * sai + gk * sak] : 0.0;
      const int gj = j0 + ii;
      Bs[kk][ii] = (gj < N && gk < K) ? B[gk * sbk + gj * sbj] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = As[kk][tx + 16 * q];
        b[q] = Bs[kk][ty + 16 * q];
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int gi = i0 + tx + 16 * x, gj = j0 + ty + 16 * y;
      if (gi < M && gj < N) C[(size_t)gj * ldc + gi] = acc[x][y];
    }
}

static void gemm(snmfnat_ctx* ctx, int M, int N, int K, const double* A, long sai, long sak, const double* B, long sbk,
                 long sbj, double* C, int ldc) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((M + GT - 1) / GT, (N + GT - 1) / GT);
  gemm_f64_kernel<<<grid, 256, 0, ctx->stream>>>(M, N, K, A, sai, sak, B, sbk, sbj, C, ldc);
  count_launch(ctx);
}

// ------------------------------------------------------------------------------------------------- elementwise
static inline int ew_grid(snmfnat_ctx* ctx, size_t n) {
  size_t b = (n + 255) / 256;
  const size_t cap = (size_t)ctx->sm_count * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
#define EW_LOOP(n) for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

__global__ void k_floor(double* x, size_t n, double flr) { EW_LOOP(n) x[i] = fmax(x[i], flr); }
__global__ void k_mask_floor(double* v, const double* mk, size_t n, double flr) { EW_LOOP(n) v[i] = fmax(v[i] * mk[i], flr); }
// numerator / denominator operands of the MU steps (sparse_nmf.m:190-206 / 213-240)
__global__ void k_ratio(const double* v, const double* lam, double* t1, double* t2, size_t n, double beta) {
  EW_LOOP(n) {
    const double l = lam[i], x = v[i];
    if (beta == 1.0) t1[i] = x / l;                       // v./lambda
    else if (beta == 2.0) { t1[i] = x; t2[i] = l; }       // v , lambda
    else { t1[i] = x * pow(l, beta - 2.0); t2[i] = pow(l, beta - 1.0); }
  }
}
__global__ void k_impute(double* v, const double* mk, const double* est, size_t n, double flr, int soft) {
  EW_LOOP(n) {
    const double m = mk[i];
    const double inv = soft ? (1.0 - m) : (m == 0.0 ? 1.0 : 0.0);
    v[i] = fmax(v[i] * m + fmax(est[i], flr) * inv, flr);   // snmf_mdi.m:251-254
  }
}
// column sums of an M x N column-major matrix (one block per column, fixed order)
__global__ void k_colsum(const double* A, int M, int N, int lda, double* out, int square) {
  __shared__ double sc[32];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double x = A[(size_t)j * lda + i];
    s += square ? x * x : x;
  }
  s = block_sum(s, sc);
  if (threadIdx.x == 0) out[j] = s;
}
// row sums of an M x N column-major matrix (one thread per row)
__global__ void k_rowsum(const double* A, int M, int N, int lda, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double s = 0.0;
  for (int j = 0; j < N; ++j) s += A[(size_t)j * lda + i];
  out[i] = s;
}
// column sums of the elementwise product of two M x N matrices
__global__ void k_colsum_prod(const double* A, const double* B, int M, int N, double* out) {
  __shared__ double sc[32];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) s = fma(A[(size_t)j * M + i], B[(size_t)j * M + i], s);
  s = block_sum(s, sc);
  if (threadIdx.x == 0) out[j] = s;
}
// h(ind,:) = h(ind,:) .* dmh ./ max(dph, flr)      dph = base (k) [+ dphm (k x n)] + sparsity(ind,:)
__global__ void k_update_h(double* h, int r, int n, const int* ind, int k, const double* dmh, const double* colsum_w,
                           const double* dphm, const double* sp, int sp_rows, int sp_cols, double flr) {
  EW_LOOP((size_t)k * n) {
    const int a = (int)(i % k), t = (int)(i / k);
    const int row = ind[a];
    double spv = sp[(sp_rows == 1 ? 0 : row) + (size_t)(sp_cols == 1 ? 0 : t) * sp_rows];
    double d = (dphm ? dphm[i] : colsum_w[a]) + spv;
    d = fmax(d, flr);
    double* hp = h + (size_t)t * r + row;
    *hp = *hp * dmh[i] / d;
  }
}
// w(:,ind) = w(:,ind) .* dmw ./ max(dpw, flr)  (sparse_nmf.m:215-239)
//   dpw = P + cs_dp(a) * w ; dmw = Q + cs_dm(a) * w   with P / Q either a per-column scalar (beta==1: hs) or an F x k matrix
__global__ void k_update_w(double* w, int F, const int* ind, int k, const double* Pm, const double* Pvec,
                           const double* cs_dp, const double* Qm, const double* cs_dm, double flr) {
  EW_LOOP((size_t)F * k) {
    const int f = (int)(i % F), a = (int)(i / F);
    double* wp = w + (size_t)ind[a] * F + f;
    const double wv = *wp;
    const double dpw = fmax((Pm ? Pm[i] : Pvec[a]) + cs_dp[a] * wv, flr);
    const double dmw = Qm[i] + cs_dm[a] * wv;
    *wp = wv * dmw / dpw;
  }
}
__global__ void k_scale_cols(double* A, int M, int N, const double* nrm2, int divide) {
  EW_LOOP((size_t)M * N) {
    const int j = (int)(i / M);
    const double s = sqrt(nrm2[j]);
    A[i] = divide ? A[i] / s : A[i] * s;
  }
}
__global__ void k_scale_rows(double* H, int r, int n, const double* nrm2) {
  EW_LOOP((size_t)r * n) H[i] = H[i] * sqrt(nrm2[i % r]);
}
__global__ void k_gather_cols(const double* w, int F, const int* ind, int k, double* out) {
  EW_LOOP((size_t)F * k) out[i] = w[(size_t)ind[i / F] * F + (i % F)];
}
__global__ void k_gather_rows(const double* h, int r, int n, const int* ind, int k, double* out) {
  EW_LOOP((size_t)k * n) out[i] = h[(size_t)(i / k) * r + ind[i % k]];
}
__global__ void k_mul(const double* a, const double* b, double* o, size_t n) { EW_LOOP(n) o[i] = a[i] * b[i]; }

// divergence terms (sparse_nmf.m:248-258) -> per-block partials, then a fixed-order final sum
__global__ void k_div_partial(const double* v, const double* lam, size_t n, double beta, double* part) {
  __shared__ double sc[32];
  double s = 0.0;
  EW_LOOP(n) {
    const double x = v[i], l = lam[i];
    if (beta == 1.0) s += x * log(x / l) - x + l;
    else if (beta == 2.0) s += (x - l) * (x - l);
    else if (beta == 0.0) s += x / l - log(x / l) - 1.0;
    else s += (pow(x, beta) + (beta - 1.0) * pow(l, beta) - beta * x * pow(l, beta - 1.0)) / (beta * (beta - 1.0));
  }
  s = block_sum(s, sc);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void k_sph_partial(const double* h, int r, int n, const double* sp, int sp_rows, int sp_cols, double* part) {
  __shared__ double sc[32];
  double s = 0.0;
  EW_LOOP((size_t)r * n) {
    const int row = (int)(i % r), t = (int)(i / r);
    s += sp[(sp_rows == 1 ? 0 : row) + (size_t)(sp_cols == 1 ? 0 : t) * sp_rows] * h[i];
  }
  s = block_sum(s, sc);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void k_final_sum(const double* part, int n, double* out) {
  __shared__ double sc[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = block_sum(s, sc);
  if (threadIdx.x == 0) *out = s;
}
// snmf_mdi.m:297-303: v_MDI = max(v.*Dm + (Nt .* v_est) .* inv, flr)
__global__ void k_mdi_final(const double* v, const double* mk, const double* est, const double* num, const double* den,
                            int F, int n, double flr, int soft, double* out) {
  EW_LOOP((size_t)F * n) {
    const int t = (int)(i / F);
    const double m = mk[i];
    const double inv = soft ? (1.0 - m) : (m == 0.0 ? 1.0 : 0.0);
    const double Nt = num[t] / fmax(den[t], flr);
    out[i] = fmax(v[i] * m + (Nt * fmax(est[i], flr)) * inv, flr);
  }
}

// ------------------------------------------------------------------------------------------------- driver
struct NmfResult {
  int iters = 0;
};

// All pointers are device pointers except sparsity_host / index vectors / outputs.
static void sparse_nmf_device(snmfnat_ctx* ctx, int F, int n, int r, const snmfnat_nmf_opts& o, const double* sp_host,
                              const uint8_t* w_ind_h, const uint8_t* h_ind_h, double* dv, double* dw, double* dh,
                              const double* dmask, int soft, double* div_out, double* cost_out, int* iters_out) {
  cudaStream_t st = ctx->stream;
  const double flr = 1e-9;  // sparse_nmf.m:166
  double beta;
  switch (o.cf) {
    case SNMFNAT_CF_IS: beta = 0.0; break;
    case SNMFNAT_CF_KL: beta = 1.0; break;
    case SNMFNAT_CF_ED: beta = 2.0; break;
    default: beta = o.beta_div; break;
  }
  const int sp_rows = o.sparsity_rows <= 1 ? 1 : o.sparsity_rows, sp_cols = o.sparsity_cols <= 1 ? 1 : o.sparsity_cols;
  SN_REQUIRE((sp_rows == 1 || sp_rows == r) && (sp_cols == 1 || sp_cols == n), SNMFNAT_EINVAL,
             "sparsity must be 1x1, r x 1 or r x n");
  std::vector<int> hidx, widx;
  for (int i = 0; i < r; ++i) {
    if (!h_ind_h || h_ind_h[i]) hidx.push_back(i);
    if (!w_ind_h || w_ind_h[i]) widx.push_back(i);
  }
  const int kh = (int)hidx.size(), kw = (int)widx.size();
  const size_t Fn = (size_t)F * n;
  DevBuf<double> lam, t1, t2, sp, wsel, hsel, dmh, dphm, G1, G2, vec, part, scal;
  DevBuf<int> d_hidx, d_widx;
  lam.alloc(Fn); t1.alloc(Fn); t2.alloc(Fn);
  sp.alloc((size_t)sp_rows * sp_cols);
  SN_CUDA(cudaMemcpyAsync(sp.p, sp_host, sp.n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (kh) { d_hidx.alloc(kh); SN_CUDA(cudaMemcpyAsync(d_hidx.p, hidx.data(), kh * sizeof(int), cudaMemcpyHostToDevice, st)); }
  if (kw) { d_widx.alloc(kw); SN_CUDA(cudaMemcpyAsync(d_widx.p, widx.data(), kw * sizeof(int), cudaMemcpyHostToDevice, st)); }
  const int kmax = kh > kw ? kh : kw;
  wsel.alloc((size_t)F * (kmax > 0 ? kmax : 1));
  hsel.alloc((size_t)(kw > 0 ? kw : 1) * n);
  dmh.alloc((size_t)(kh > 0 ? kh : 1) * n);
  dphm.alloc((size_t)(kh > 0 ? kh : 1) * n);
  G1.alloc((size_t)F * (kw > 0 ? kw : 1));
  G2.alloc((size_t)F * (kw > 0 ? kw : 1));
  vec.alloc((size_t)4 * (r > n ? r : n) + 16);
  const int nblk = ew_grid(ctx, Fn);
  part.alloc((size_t)2 * nblk + 2 * ew_grid(ctx, (size_t)r * n) + 8);
  scal.alloc(4);
  auto launches = [&](int k) { count_launch(ctx, k); };

  // normalise W, rescale H (sparse_nmf.m:157-160)
  double* nrm2 = vec.p;  // [r]
  k_colsum<<<r, 128, 0, st>>>(dw, F, r, F, nrm2, 1);
  k_scale_cols<<<ew_grid(ctx, (size_t)F * r), 256, 0, st>>>(dw, F, r, nrm2, 1);
  k_scale_rows<<<ew_grid(ctx, (size_t)r * n), 256, 0, st>>>(dh, r, n, nrm2);
  launches(3);
  auto update_lambda = [&]() {  // lambda = max(w*h, flr)
    gemm(ctx, F, n, r, dw, 1, F, dh, 1, r, lam.p, F);
    k_floor<<<nblk, 256, 0, st>>>(lam.p, Fn, flr);
    launches(1);
  };
  update_lambda();
  if (dmask) k_mask_floor<<<nblk, 256, 0, st>>>(dv, dmask, Fn, flr);   // snmf_mdi.m:175
  else k_floor<<<nblk, 256, 0, st>>>(dv, Fn, flr);                     // sparse_nmf.m:169
  launches(1);

  double last_cost = INFINITY;
  int its = o.max_iter;
  if (div_out) std::memset(div_out, 0, sizeof(double) * (o.max_iter > 0 ? o.max_iter : 0));
  if (cost_out) std::memset(cost_out, 0, sizeof(double) * (o.max_iter > 0 ? o.max_iter : 0));
  for (int it = 1; it <= o.max_iter; ++it) {
    if (kh > 0) {  // ---- H update (:189-208)
      k_ratio<<<nblk, 256, 0, st>>>(dv, lam.p, t1.p, t2.p, Fn, beta);
      k_gather_cols<<<ew_grid(ctx, (size_t)F * kh), 256, 0, st>>>(dw, F, d_hidx.p, kh, wsel.p);
      launches(2);
      gemm(ctx, kh, n, F, wsel.p, F, 1, t1.p, 1, F, dmh.p, kh);              // w(:,ind)' * T1
      double* csw = vec.p + r;
      if (beta == 1.0) {
        k_colsum<<<kh, 128, 0, st>>>(wsel.p, F, kh, F, csw, 0);
        launches(1);
      } else {
        gemm(ctx, kh, n, F, wsel.p, F, 1, t2.p, 1, F, dphm.p, kh);           // w(:,ind)' * T2
      }
      k_update_h<<<ew_grid(ctx, (size_t)kh * n), 256, 0, st>>>(dh, r, n, d_hidx.p, kh, dmh.p, csw,
                                                                beta == 1.0 ? nullptr : dphm.p, sp.p, sp_rows, sp_cols, flr);
      launches(1);
      update_lambda();
    }
    if (kw > 0) {  // ---- W update (:212-244)
      k_ratio<<<nblk, 256, 0, st>>>(dv, lam.p, t1.p, t2.p, Fn, beta);
      k_gather_rows<<<ew_grid(ctx, (size_t)kw * n), 256, 0, st>>>(dh, r, n, d_widx.p, kw, hsel.p);
      k_gather_cols<<<ew_grid(ctx, (size_t)F * kw), 256, 0, st>>>(dw, F, d_widx.p, kw, wsel.p);
      launches(3);
      gemm(ctx, F, kw, n, t1.p, 1, F, hsel.p, kw, 1, G1.p, F);               // T1 * h(ind,:)'
      double* c1 = vec.p + r;          // [kw]
      double* c2 = vec.p + 2 * r;      // [kw]
      double* hs = vec.p + 3 * r;      // [kw]
      if (beta == 1.0) {
        // dpw = hs + colsum(G1.*w).*w ; dmw = G1 + (hs .* colsum(w)).*w
        k_rowsum<<<(kw + 127) / 128, 128, 0, st>>>(hsel.p, kw, n, kw, hs);
        k_colsum_prod<<<kw, 128, 0, st>>>(G1.p, wsel.p, F, kw, c1);
        k_colsum<<<kw, 128, 0, st>>>(wsel.p, F, kw, F, c2, 0);
        k_mul<<<1, 256, 0, st>>>(c2, hs, c2, kw);
        k_update_w<<<ew_grid(ctx, (size_t)F * kw), 256, 0, st>>>(dw, F, d_widx.p, kw, nullptr, hs, c1, G1.p, c2, flr);
        launches(5);
      } else {
        // beta==2: dpw = lambda*h' + colsum(v*h'.*w).*w ; dmw = v*h' + colsum(lambda*h'.*w).*w  (T1 = v-term, T2 = lambda-term)
        gemm(ctx, F, kw, n, t2.p, 1, F, hsel.p, kw, 1, G2.p, F);
        k_colsum_prod<<<kw, 128, 0, st>>>(G1.p, wsel.p, F, kw, c1);
        k_colsum_prod<<<kw, 128, 0, st>>>(G2.p, wsel.p, F, kw, c2);
        k_update_w<<<ew_grid(ctx, (size_t)F * kw), 256, 0, st>>>(dw, F, d_widx.p, kw, G2.p, nullptr, c1, G1.p, c2, flr);
        launches(3);
      }
      k_colsum<<<r, 128, 0, st>>>(dw, F, r, F, nrm2, 1);
      k_scale_cols<<<ew_grid(ctx, (size_t)F * r), 256, 0, st>>>(dw, F, r, nrm2, 1);   // :242 (all columns)
      launches(2);
      update_lambda();
    }
    if (dmask) {  // snmf_mdi.m:251-255 (v_est = max(w*h, flr) == lambda)
      k_impute<<<nblk, 256, 0, st>>>(dv, dmask, lam.p, Fn, flr, soft);
      launches(1);
    }
    if (o.cost_check) {  // :247-285
      const int nb2 = ew_grid(ctx, (size_t)r * n);
      k_div_partial<<<nblk, 256, 0, st>>>(dv, lam.p, Fn, beta, part.p);
      k_final_sum<<<1, 256, 0, st>>>(part.p, nblk, scal.p);
      k_sph_partial<<<nb2, 256, 0, st>>>(dh, r, n, sp.p, sp_rows, sp_cols, part.p + nblk);
      k_final_sum<<<1, 256, 0, st>>>(part.p + nblk, nb2, scal.p + 1);
      launches(4);
      double hc[2];
      SN_CUDA(cudaMemcpyAsync(hc, scal.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      SN_CUDA(cudaStreamSynchronize(st));
      const double div = hc[0], cost = hc[0] + hc[1];
      if (div_out) div_out[it - 1] = div;
      if (cost_out) cost_out[it - 1] = cost;
      if (it > 1 && o.conv_eps > 0) {
        const double e = std::fabs(cost - last_cost) / last_cost;
        if (e < o.conv_eps) {
          its = it;
          break;
        }
      }
      last_cost = cost;
    }
  }
  if (iters_out) *iters_out = its;
  check_launch(ctx, "sparse_nmf (generic)");
}

static void validate_nmf_args(snmfnat_ctx* ctx, const double* v, int F, int n, int r, const snmfnat_nmf_opts* o,
                              const double* sp, const double* w0, const double* h0) {
  SN_REQUIRE(ctx && v && o && sp, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(F > 0 && n > 0 && r > 0 && o->max_iter >= 0, SNMFNAT_EINVAL, "bad dimensions");
  SN_REQUIRE(w0 != nullptr, SNMFNAT_EINVAL, "init_w is required: random initialisation stays with the caller "
                                           "(sparse_nmf.m:116-131 draws it with rand)");
  SN_REQUIRE(h0 != nullptr, SNMFNAT_EINVAL, "init_h is required: random initialisation stays with the caller "
                                           "(sparse_nmf.m:133-134 draws it with rand)");
  SN_REQUIRE(o->precision == 0, SNMFNAT_EUNSUPPORTED, "snmfnat_sparse_nmf computes in float64; use snmfnat_train_* for TF32");
  SN_CUDA(cudaSetDevice(ctx->device));
}

}  // namespace snmfnat

using namespace snmfnat;

extern "C" {

int snmfnat_sparse_nmf(snmfnat_ctx* ctx, const double* v, int F, int n, int r, const snmfnat_nmf_opts* opts,
                       const double* sparsity, const double* init_w, const double* init_h, const uint8_t* w_ind,
                       const uint8_t* h_ind, double* w, double* h, double* div, double* cost, int* iters) {
  SN_API_BEGIN
  validate_nmf_args(ctx, v, F, n, r, opts, sparsity, init_w, init_h);
  SN_REQUIRE(w && h, SNMFNAT_EINVAL, "NULL output");
  DevBuf<double> dv, dw, dh;
  dv.alloc((size_t)F * n); dw.alloc((size_t)F * r); dh.alloc((size_t)r * n);
  cudaStream_t st = ctx->stream;
  SN_CUDA(cudaMemcpyAsync(dv.p, v, dv.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dw.p, init_w, dw.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dh.p, init_h, dh.n * sizeof(double), cudaMemcpyHostToDevice, st));
  sparse_nmf_device(ctx, F, n, r, *opts, sparsity, w_ind, h_ind, dv.p, dw.p, dh.p, nullptr, 0, div, cost, iters);
  SN_CUDA(cudaMemcpyAsync(w, dw.p, dw.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaMemcpyAsync(h, dh.p, dh.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  SN_API_END
}

int snmfnat_snmf_mdi(snmfnat_ctx* ctx, const double* v, const double* mask, int soft, int F, int n, int r,
                     const snmfnat_nmf_opts* opts, const double* sparsity, const double* init_w, const double* init_h,
                     const uint8_t* w_ind, const uint8_t* h_ind, double* v_mdi, double* h, double* div, double* cost,
                     int* iters) {
  SN_API_BEGIN
  validate_nmf_args(ctx, v, F, n, r, opts, sparsity, init_w, init_h);
  SN_REQUIRE(mask && v_mdi && h, SNMFNAT_EINVAL, "NULL argument");
  const size_t Fn = (size_t)F * n;
  DevBuf<double> dv, dw, dh, dm, est, tmp, num, den, out;
  dv.alloc(Fn); dw.alloc((size_t)F * r); dh.alloc((size_t)r * n); dm.alloc(Fn); est.alloc(Fn); tmp.alloc(Fn);
  num.alloc(n); den.alloc(n); out.alloc(Fn);
  cudaStream_t st = ctx->stream;
  SN_CUDA(cudaMemcpyAsync(dv.p, v, Fn * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dm.p, mask, Fn * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dw.p, init_w, dw.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dh.p, init_h, dh.n * sizeof(double), cudaMemcpyHostToDevice, st));
  sparse_nmf_device(ctx, F, n, r, *opts, sparsity, w_ind, h_ind, dv.p, dw.p, dh.p, dm.p, soft, div, cost, iters);
  // gain-matched imputation (snmf_mdi.m:297-303)
  gemm(ctx, F, n, r, dw.p, 1, F, dh.p, 1, r, est.p, F);
  const int g = ew_grid(ctx, Fn);
  k_floor<<<g, 256, 0, st>>>(est.p, Fn, 1e-9);
  k_mul<<<g, 256, 0, st>>>(dv.p, dm.p, tmp.p, Fn);
  k_colsum<<<n, 128, 0, st>>>(tmp.p, F, n, F, num.p, 0);
  k_mul<<<g, 256, 0, st>>>(est.p, dm.p, tmp.p, Fn);
  k_colsum<<<n, 128, 0, st>>>(tmp.p, F, n, F, den.p, 0);
  k_mdi_final<<<g, 256, 0, st>>>(dv.p, dm.p, est.p, num.p, den.p, F, n, 1e-9, soft, out.p);
  count_launch(ctx, 6);
  check_launch(ctx, "snmf_mdi");
  SN_CUDA(cudaMemcpyAsync(v_mdi, out.p, Fn * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaMemcpyAsync(h, dh.p, dh.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  SN_API_END
}

int snmfnat_dnmf_adapt(snmfnat_ctx* ctx, const double* Y, const double* D, const double* B, int F, int n, int R_x,
                       int R_d, const snmfnat_nmf_opts* opts, const double* sparsity, const double* h_init,
                       double* B_a) {
  SN_API_BEGIN
  const int r = R_x + R_d;
  validate_nmf_args(ctx, Y, F, n, r, opts, sparsity, B, h_init);
  SN_REQUIRE(D && B_a && R_x >= 0 && R_d > 0, SNMFNAT_EINVAL, "bad argument");
  SN_REQUIRE(opts->sparsity_rows <= 1 && opts->sparsity_cols <= 1, SNMFNAT_EUNSUPPORTED,
             "DNMF_adapt supports a scalar sparsity (the two inner solves have different ranks)");
  cudaStream_t st = ctx->stream;
  DevBuf<double> dv, dw, dh, dv2, dw2, dh2;
  dv.alloc((size_t)F * n); dw.alloc((size_t)F * r); dh.alloc((size_t)r * n);
  SN_CUDA(cudaMemcpyAsync(dv.p, Y, dv.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dw.p, B, dw.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dh.p, h_init, dh.n * sizeof(double), cudaMemcpyHostToDevice, st));
  std::vector<uint8_t> zeros(r, 0), ones(r, 1);
  // A_hat: H-solve on Y with the full basis (DNMF_adapt.m:3-7)
  sparse_nmf_device(ctx, F, n, r, *opts, sparsity, zeros.data(), ones.data(), dv.p, dw.p, dh.p, nullptr, 0, nullptr,
                    nullptr, nullptr);
  // W-only update of the noise basis on D with H fixed (:15-20); init_w is the ORIGINAL B(:,R_x+1:end)
  dv2.alloc((size_t)F * n); dw2.alloc((size_t)F * R_d); dh2.alloc((size_t)R_d * n);
  SN_CUDA(cudaMemcpyAsync(dv2.p, D, dv2.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dw2.p, B + (size_t)R_x * F, dw2.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpy2DAsync(dh2.p, (size_t)R_d * sizeof(double), dh.p + R_x, (size_t)r * sizeof(double),
                            (size_t)R_d * sizeof(double), n, cudaMemcpyDeviceToDevice, st));
  sparse_nmf_device(ctx, F, n, R_d, *opts, sparsity, ones.data(), zeros.data(), dv2.p, dw2.p, dh2.p, nullptr, 0, nullptr,
                    nullptr, nullptr);
  SN_CUDA(cudaMemcpyAsync(B_a, dw2.p, dw2.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  SN_API_END
}

}  // extern "C"
