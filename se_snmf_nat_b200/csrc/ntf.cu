// GIST_NTF / GIST_NTF_C (src/GIST_NTF.m:1-160, src/GIST_NTF_C.m): KL non-negative tensor factorisation of a
// multi-channel magnitude tensor  S(h,n,m) ~ sum_k C(h,k) B(n,k) A(m,k)  in which only the channel gains C are updated
// (C_UPDATE = 1, A_UPDATE = 0, A = ones(M,K), GIST_NTF.m:4-6,15).  The Khatri-Rao products of the reference
// (src/kr.m) are never materialised: X_hat and the two C-update sums are computed directly from the factors.
#include <algorithm>
#include <cmath>
#include <vector>
#include "common.cuh"

namespace snmfnat {

// Bn = sqrt(sum(B.^2)); B = B ./ Bn; C = C .* Bn   (GIST_NTF.m:27-29).  One block per atom.
__global__ void ntf_norm_kernel(double* __restrict__ B, int N, int K, double* __restrict__ C, int Ch, double* __restrict__ bsum) {
  __shared__ double scratch[40];
  const int k = blockIdx.x;
  double s = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s = fma(B[(size_t)k * N + n], B[(size_t)k * N + n], s);
  const double bn = sqrt(block_sum(s, scratch));
  double t = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double x = B[(size_t)k * N + n] / bn;
    B[(size_t)k * N + n] = x;
    t += x;
  }
  t = block_sum(t, scratch);
  if (threadIdx.x == 0) bsum[k] = t;   // sum_n B(n,k) of the normalised dictionary (the O-weighted sum of :112)
  for (int h = threadIdx.x; h < Ch; h += blockDim.x) C[(size_t)k * Ch + h] *= bn;
}

// X_hat(h,:,m) = max(sum_k C(h,k) B(:,k) A(m,k), flr); P = max(S ./ X_hat, flr)  (:40-43,126-129) and the KL terms
// of this slice (:131).  Grid (M, Ch).  S is Ch x N x M column-major; P is stored [h][m][n].
__global__ void ntf_xhat_kernel(const double* __restrict__ S, const double* __restrict__ B, const double* __restrict__ C,
                                const double* __restrict__ A, int Ch, int N, int M, int K, double flr,
                                double* __restrict__ P, double* __restrict__ div_part) {
  extern __shared__ double ck[];   // [K] C(h,k) A(m,k)
  __shared__ double scratch[40];
  const int m = blockIdx.x, h = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) ck[k] = C[(size_t)k * Ch + h] * (A ? A[(size_t)k * M + m] : 1.0);
  __syncthreads();
  double dsum = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    double x = 0.0;
    for (int k = 0; k < K; ++k) x = fma(ck[k], B[(size_t)k * N + n], x);
    x = fmax(x, flr);
    const double s = S[((size_t)m * N + n) * Ch + h];
    P[((size_t)h * M + m) * N + n] = fmax(s / x, flr);
    dsum += s * log(s / x) - s + x;
  }
  dsum = block_sum(dsum, scratch);
  if (threadIdx.x == 0) div_part[(size_t)h * M + m] = dsum;
}

// C(h,k) = max(C .* max(PBA, flr) ./ (max(OBA, flr) + sparsity), flr) with PBA(h,k) = sum_{n,m} P(h,n,m) B(n,k) A(m,k),
// OBA(h,k) = sum_{n,m} B(n,k) A(m,k)   (:96-116).  Grid (K, Ch).
__global__ void ntf_cupdate_kernel(const double* __restrict__ P, const double* __restrict__ B, const double* __restrict__ A,
                                   const double* __restrict__ bsum, int Ch, int N, int M, int K, double flr, double sparsity,
                                   const double* __restrict__ Cin, double* __restrict__ Cout) {
  __shared__ double scratch[40];
  const int k = blockIdx.x, h = blockIdx.y;
  double acc = 0.0, asum = 0.0;
  for (int m = 0; m < M; ++m) {
    const double a = A ? A[(size_t)k * M + m] : 1.0;
    const double* p = P + ((size_t)h * M + m) * N;
    double s = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) s = fma(p[n], B[(size_t)k * N + n], s);
    acc = fma(a, s, acc);
    asum += a;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    const double pba = fmax(acc, flr), oba = fmax(bsum[k] * asum, flr);
    const double c = Cin[(size_t)k * Ch + h];
    Cout[(size_t)k * Ch + h] = fmax(c * pba / (oba + sparsity), flr);
  }
}

__global__ void ntf_sum_kernel(const double* __restrict__ x, int n, double* __restrict__ out) {
  __shared__ double scratch[40];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s;
}

}  // namespace snmfnat

extern "C" int snmfnat_gist_ntf(snmfnat_ctx* ctx, const double* S_mag, int Channel, int N, int M, const double* B, int K,
                                const double* C_init, const double* A, double sparsity, double flr, int max_iter,
                                double conv_eps, int cost_check, double* C_out, double* div, double* cost, int* iters) {
  using namespace snmfnat;
  SN_API_BEGIN
  SN_REQUIRE(ctx && S_mag && B && C_init && C_out, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(Channel > 0 && N > 0 && M > 0 && K > 0 && max_iter >= 0, SNMFNAT_EINVAL, "bad size");
  SN_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevBuf<double> dS, dB, dC[2], dA, dP, dpart, dbsum, dscal;
  dS.alloc((size_t)Channel * N * M); dB.alloc((size_t)N * K); dC[0].alloc((size_t)Channel * K); dC[1].alloc((size_t)Channel * K);
  dP.alloc((size_t)Channel * N * M); dpart.alloc((size_t)Channel * M); dbsum.alloc(K); dscal.alloc(2);
  SN_CUDA(cudaMemcpyAsync(dS.p, S_mag, dS.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dB.p, B, dB.n * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dC[0].p, C_init, dC[0].n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (A) {
    dA.alloc((size_t)M * K);
    SN_CUDA(cudaMemcpyAsync(dA.p, A, dA.n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  ntf_norm_kernel<<<K, 256, 0, st>>>(dB.p, N, K, dC[0].p, Channel, dbsum.p);
  count_launch(ctx);
  int cur = 0;
  auto xhat = [&]() {
    ntf_xhat_kernel<<<dim3(M, Channel), 256, (size_t)K * sizeof(double), st>>>(dS.p, dB.p, dC[cur].p, A ? dA.p : nullptr, Channel, N,
                                                                                 M, K, flr, dP.p, dpart.p);
    count_launch(ctx);
  };
  xhat();
  check_launch(ctx, "ntf_xhat_kernel");
  int done = 0;
  double last_cost = 0.0;
  for (int i = 1; i <= max_iter; ++i) {
    ntf_cupdate_kernel<<<dim3(K, Channel), 256, 0, st>>>(dP.p, dB.p, A ? dA.p : nullptr, dbsum.p, Channel, N, M, K, flr, sparsity,
                                                          dC[cur].p, dC[cur ^ 1].p);
    count_launch(ctx);
    cur ^= 1;
    xhat();
    done = i;
    // GIST_NTF.m computes the objective on every iteration (:131-134); GIST_NTF_C.m only when p.cost_check
    if (cost_check != 0) {
      ntf_sum_kernel<<<1, 256, 0, st>>>(dpart.p, Channel * M, dscal.p);
      ntf_sum_kernel<<<1, 256, 0, st>>>(dC[cur].p, Channel * K, dscal.p + 1);
      count_launch(ctx, 2);
      double hs[2];
      SN_CUDA(cudaMemcpyAsync(hs, dscal.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      SN_CUDA(cudaStreamSynchronize(st));
      const double c = hs[0] + sparsity * hs[1];
      if (div) div[i - 1] = hs[0];
      if (cost) cost[i - 1] = c;
      if (i > 1 && conv_eps > 0.0 && std::fabs(c - last_cost) / last_cost < conv_eps) break;   // :143-154
      last_cost = c;
    }
  }
  check_launch(ctx, "GIST_NTF iteration");
  SN_CUDA(cudaMemcpyAsync(C_out, dC[cur].p, dC[cur].n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  if (iters) *iters = done;
  SN_API_END
}
