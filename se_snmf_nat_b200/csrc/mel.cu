// Mel-domain separation mode (p.B_sep_mode = 'Mel'): the triangular filterbank of src/mel_matrix.m, the DFT -> Mel
// projection of the separation input (src/bnmf_sep_event_RT_IS16.m:107-119), the Mel -> DFT conversion of the separated
// spectra (:165-172,187-195,205-211) and the Mel image of the noise history the adaptation solves on (:295-301).
// The H-solve and W-solve themselves are the ordinary kernels run on a Mel-sized view of the slot state.
#include <algorithm>
#include <cmath>
#include <vector>
#include "online.cuh"

namespace snmfnat {

// round() of MATLAB: halves away from zero
static double mround(double x) { return x < 0 ? -std::floor(-x + 0.5) : std::floor(x + 0.5); }

// [M] = mel_matrix(fs, NbCh, Nfft, warp, fhigh), src/mel_matrix.m:9-38.  M is (Nfft/2+1) x NbCh, column-major.
void mel_matrix_host(int fs, int NbCh, int Nfft, double warp, double fhigh, double* M) {
  const int rows_out = Nfft / 2 + 1;
  const double low = 2595.0 * std::log10(1.0 + 64.0 / 700.0);
  const double nyq = 2595.0 * std::log10(1.0 + fhigh / 700.0);
  std::vector<long> start(NbCh), end(NbCh);
  for (int k = 0; k < NbCh; ++k) {
    const double sm = low + (double)k / (NbCh + 1) * (nyq - low);
    const double fcen = warp * 700.0 * (std::pow(10.0, sm / 2595.0) - 1.0);
    start[k] = (long)mround((double)Nfft / fs * fcen) + 1;                                   // 1-based bins
    const double em = low + (double)(k + 2) / (NbCh + 1) * (nyq - low);
    end[k] = (long)mround(warp * Nfft / fs * 700.0 * (std::pow(10.0, em / 2595.0) - 1.0)) + 1;
  }
  for (size_t i = 0; i < (size_t)rows_out * NbCh; ++i) M[i] = 0.0;
  for (int k = 0; k < NbCh; ++k) {
    const long tot = end[k] - start[k] + 1;
    const long next_start = (k + 1 < NbCh) ? start[k + 1] : end[NbCh - 2];
    const long lowlen = next_start - start[k] + 1;
    const long hilen = tot - lowlen + 1;
    // rising edge then falling edge; the falling edge is assigned second and wins on the shared bin
    for (long i = 1; i <= lowlen; ++i) {
      const long row = start[k] + i - 1;
      if (row >= 1 && row <= rows_out) M[(size_t)k * rows_out + (row - 1)] = (double)i / (double)lowlen;
    }
    for (long i = 0; i < hilen; ++i) {
      const long row = end[k] - hilen + 1 + i;
      if (row >= 1 && row <= rows_out) M[(size_t)k * rows_out + (row - 1)] = (double)(hilen - i) / (double)hilen;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Y_sep = (melmat * Ym ./ ||melmat * Ym||_2 + 1e-9) * ||Ym||_2 for every frame (:107-119).  One CTA per frame.
__global__ void mel_project_kernel(const double* __restrict__ M, int n1, int LD1, int F, int LDF,
                                   const double* __restrict__ Ym, long long NF, double* __restrict__ Ysep) {
  __shared__ double band[256];
  __shared__ double scratch[40];
  const long long frame = blockIdx.x;
  if (frame >= NF) return;
  const double* y = Ym + (size_t)frame * LDF;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int b = warp; b < n1; b += nw) {
    const double* m = M + (size_t)b * LDF;
    double s = 0.0;
    for (int f = lane; f < F; f += 32) s = fma(m[f], y[f], s);
    s = warp_sum(s);
    if (lane == 0) band[b] = s;
  }
  double t2 = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) t2 = fma(y[f], y[f], t2);
  t2 = block_sum(t2, scratch);   // barriers inside: band[] is complete afterwards
  double v2 = 0.0;
  for (int b = threadIdx.x; b < n1; b += blockDim.x) v2 = fma(band[b], band[b], v2);
  v2 = block_sum(v2, scratch);
  const double vn = sqrt(v2), tn = sqrt(t2);
  double* o = Ysep + (size_t)frame * LD1;
  for (int b = threadIdx.x; b < LD1; b += blockDim.x) o[b] = (b < n1) ? (band[b] / vn + 1e-9) * tn : 0.0;
}

void launch_mel_project(snmfnat_ctx* ctx, const double* M, int n1, int LD1, int F, int LDF, const double* Ym, long long NF,
                        double* Ysep) {
  if (NF <= 0) return;
  SN_REQUIRE(n1 <= 256, SNMFNAT_EUNSUPPORTED, "Mel mode supports up to 256 bands (got %d)", n1);
  mel_project_kernel<<<(unsigned)NF, 256, 0, ctx->stream>>>(M, n1, LD1, F, LDF, Ym, NF, Ysep);
  count_launch(ctx);
  check_launch(ctx, "mel_project_kernel");
}

// After the Mel-domain H-solve of a hop: X^ = melmat' * (B_Mel_x A_x), D^ = melmat' * (B_Mel_d A_d) (:165-172,187-195)
// and, on the first hop of a file, lambda_dav <- melmat' * Ym_Mel (:205-211,223-225).  One CTA per slot.
__global__ void mel_post_kernel(OnlineDims d, SlotState st, const double* __restrict__ M, int n1, int LD1,
                                const double* __restrict__ XhatM, const double* __restrict__ DhatM,
                                const double* __restrict__ Ysep, int g_step) {
  __shared__ double xm[256], dm[256], ym[256];
  const int slot = d.slot0 + (int)blockIdx.x * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;
  const long long frame = st.frame_base[slot] + g_step;
  for (int b = threadIdx.x; b < n1; b += blockDim.x) {
    xm[b] = XhatM[(size_t)slot * LD1 + b];
    dm[b] = DhatM[(size_t)slot * LD1 + b];
    ym[b] = Ysep[(size_t)frame * LD1 + b];
  }
  __syncthreads();
  const int F = d.F, LDF = d.LDF;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    double sx = 0.0, sd = 0.0, sy = 0.0;
    for (int b = 0; b < n1; ++b) {
      const double m = M[(size_t)b * LDF + f];
      sx = fma(m, xm[b], sx);
      sd = fma(m, dm[b], sd);
      sy = fma(m, ym[b], sy);
    }
    st.Xhat[(size_t)slot * LDF + f] = sx;
    st.Dhat[(size_t)slot * LDF + f] = sd;
    if (l == 1) st.lambda_dav[(size_t)slot * LDF + f] = sy;
  }
}

void launch_mel_post(snmfnat_ctx* ctx, const OnlineDims& d, const SlotState& st, const double* M, int n1, int LD1,
                     const double* XhatM, const double* DhatM, const double* Ysep, int n_active, int g_step) {
  if (n_active <= 0) return;
  mel_post_kernel<<<n_active, 256, 0, ctx->stream>>>(d, st, M, n1, LD1, XhatM, DhatM, Ysep, g_step);
  count_launch(ctx);
  check_launch(ctx, "mel_post_kernel");
}

// After the gain / gate kernel of a hop: the column the gate just appended to lambda_d_blk gets its Mel image
// (the reference recomputes melmat * lambda_d_blk for the whole history, :295-301; only one column is new).
__global__ void mel_hist_kernel(OnlineDims d, SlotState st, const double* __restrict__ M, int n1, int LD1,
                                double* __restrict__ lam_blk_mel, int g_step) {
  const int slot = d.slot0 + (int)blockIdx.x * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;
  if (!st.gated[slot]) return;
  const int m_a = d.m_a, F = d.F, LDF = d.LDF;
  const int col = (st.ring_head[slot] + m_a - 1) % m_a;   // the gate advanced the head after writing
  const double* src = st.lam_blk + ((size_t)slot * m_a + col) * LDF;
  double* dst = lam_blk_mel + ((size_t)slot * m_a + col) * LD1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int b = warp; b < n1; b += nw) {
    const double* m = M + (size_t)b * LDF;
    double s = 0.0;
    for (int f = lane; f < F; f += 32) s = fma(m[f], src[f], s);
    s = warp_sum(s);
    if (lane == 0) dst[b] = s;
  }
}

void launch_mel_hist(snmfnat_ctx* ctx, const OnlineDims& d, const SlotState& st, const double* M, int n1, int LD1,
                     double* lam_blk_mel, int n_active, int g_step) {
  if (n_active <= 0) return;
  mel_hist_kernel<<<n_active, 256, 0, ctx->stream>>>(d, st, M, n1, LD1, lam_blk_mel, g_step);
  count_launch(ctx);
  check_launch(ctx, "mel_hist_kernel");
}

// Training features (run_basis_train.m:63,70-78; run_basis_DNMF.m:15,24,33; run_basis_DNMF_Mel.m:16-27): X = S.^pow + floor
// and, when a filterbank is given, X_Mel = melmat' * X.  One CTA per frame; in/out column-major (bins x frames).
__global__ void tf_features_kernel(const double* __restrict__ S, int F, long long T, double pw, double flr,
                                   const double* __restrict__ M /* [n1][F] or null */, int n1, double* __restrict__ out) {
  extern __shared__ double xs[];
  for (long long t = blockIdx.x; t < T; t += gridDim.x) {
    const double* s = S + (size_t)t * F;
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      const double m = s[f];
      const double x = (pw == 2.0 ? m * m : (pw == 1.0 ? m : pow(m, pw))) + flr;
      if (M) xs[f] = x; else out[(size_t)t * F + f] = x;
    }
    if (!M) continue;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = warp; b < n1; b += nw) {
      const double* m = M + (size_t)b * F;
      double acc = 0.0;
      for (int f = lane; f < F; f += 32) acc = fma(m[f], xs[f], acc);
      acc = warp_sum(acc);
      if (lane == 0) out[(size_t)t * n1 + b] = acc;
    }
  }
}

}  // namespace snmfnat

extern "C" int snmfnat_tf_features(snmfnat_ctx* ctx, const double* S_mag, int F, int64_t T, double pow_, double floor_,
                                   const double* melmat, int n1, double* out) {
  using namespace snmfnat;
  SN_API_BEGIN
  SN_REQUIRE(ctx && S_mag && out && F > 0 && T >= 0, SNMFNAT_EINVAL, "bad argument");
  SN_REQUIRE(melmat == nullptr || n1 > 0, SNMFNAT_EINVAL, "n1 must be positive with a filterbank");
  SN_CUDA(cudaSetDevice(ctx->device));
  if (T == 0) return SNMFNAT_OK;
  DevBuf<double> dS, dM, dO;
  dS.alloc((size_t)F * T);
  dO.alloc((size_t)(melmat ? n1 : F) * T);
  SN_CUDA(cudaMemcpyAsync(dS.p, S_mag, dS.n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (melmat) {  // host n2 x n1 column-major == device [n1][F]
    dM.alloc((size_t)n1 * F);
    SN_CUDA(cudaMemcpyAsync(dM.p, melmat, dM.n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  const unsigned grid = (unsigned)std::min<int64_t>(T, (int64_t)ctx->sm_count * 8);
  tf_features_kernel<<<grid, 256, (size_t)F * sizeof(double), ctx->stream>>>(dS.p, F, T, pow_, floor_, melmat ? dM.p : nullptr,
                                                                             n1, dO.p);
  count_launch(ctx);
  check_launch(ctx, "tf_features_kernel");
  SN_CUDA(cudaMemcpyAsync(out, dO.p, dO.n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
  SN_API_END
}

extern "C" int snmfnat_mel_matrix(int fs, int NbCh, int Nfft, double warp, double fhigh, double* M) {
  SN_API_BEGIN
  SN_REQUIRE(M != nullptr && fs > 0 && NbCh >= 2 && Nfft >= 2 && Nfft % 2 == 0, SNMFNAT_EINVAL, "bad argument");
  snmfnat::mel_matrix_host(fs, NbCh, Nfft, warp > 0 ? warp : 1.0, fhigh > 0 ? fhigh : fs / 2.0, M);
  SN_API_END
}
