// L1 entry points that mirror the reference's stand-alone numeric functions:
//   stft_fft (src/stft_fft.m), synth_ifft_buff (src/synth_ifft_buff.m), blk_sparse (src/blk_sparse.m).
#include <cmath>
#include <vector>
#include "state.cuh"
#include "blk_sparse.cuh"

namespace snmfnat {

// ---- training STFT: frames start at samples 0, shift, 2*shift, ... while (1-based) pos < len - fftlen  (stft_fft.m:21)
__global__ void __launch_bounds__(256)
frame_signal_kernel(const double* __restrict__ s, int sz, int shift, int fftlen, double preemph,
                    const double* __restrict__ win, long long nf, double* __restrict__ frames) {
  for (long long fr = blockIdx.x; fr < nf; fr += gridDim.x) {
    const double* x = s + fr * shift;
    double* out = frames + (size_t)fr * fftlen;
    for (int i = threadIdx.x; i < fftlen; i += blockDim.x) {
      double v = 0.0;
      if (i < sz) v = win[i] * (x[i] - preemph * (i > 0 ? x[i - 1] : 0.0));   // :22-23
      out[i] = v;
    }
  }
}
// S_mag = abs(S) with DC bins := 1e-6 (:27,31), S_phase = angle(S) (:28); outputs column-major half x frame_num
__global__ void __launch_bounds__(256)
stft_fft_post_kernel(const double2* __restrict__ Y, int half, long long nf, int DCbin, double* __restrict__ mag,
                     double* __restrict__ ph) {
  const long long total = nf * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % half);
    const double2 c = Y[i];
    mag[i] = (f < DCbin) ? 0.000001 : sqrt(fma(c.x, c.x, c.y * c.y));
    ph[i] = atan2(c.y, c.x);
  }
}

// ---- synth_ifft_buff spectrum assembly (synth_ifft_buff.m:10-18)
__global__ void __launch_bounds__(256)
synth_pre_kernel(const double* __restrict__ TF_mag, const double* __restrict__ TF_phase, int freq_num, int fftlen,
                 long long nf, int DCbin_back, double pow_, double2* __restrict__ Z) {
  const int half = fftlen / 2 + 1;
  const long long total = nf * half;
  const double ip = 1.0 / pow_;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long fr = i / half;
    const int f = (int)(i - fr * half);
    auto magp = [&](int k) {
      const double m = (k < DCbin_back) ? 0.0 : TF_mag[(size_t)fr * freq_num + k];
      return (pow_ == 1.0) ? m : ((pow_ == 2.0) ? sqrt(m) : pow(m, ip));
    };
    double2 z;
    if (freq_num == fftlen) {
      // TF = TF_mag(:,i) (real, full length); real(ifft(x)) of a real x is the inverse transform of its even part
      const double a = magp(f), b = magp((fftlen - f) % fftlen);
      z = make_double2(0.5 * (a + b), 0.0);
    } else {
      const double m = magp(f);
      double sn, cs;
      sincos(TF_phase[(size_t)fr * freq_num + f], &sn, &cs);
      z = make_double2(m * cs, m * sn);
    }
    Z[i] = z;
  }
}

__global__ void __launch_bounds__(256)
blk_sparse_kernel(const double* __restrict__ X, const double* __restrict__ D, double* rb, int K, int l,
                  OnlineScalars sc, double* __restrict__ Q) {
  extern __shared__ __align__(16) double smem[];
  double* Q_s = smem;
  double* rs1 = Q_s + K;
  double* rs2 = rs1 + K;
  double* P_s = rs2 + K;
  double* scratch = P_s + K;
  // the caller's r_blk is oldest-first: slot 0 is the oldest column and receives the new one (:14)
  blk_sparse_dev(X, D, rb, K, K, l, 0, sc, Q_s, rs1, rs2, P_s, scratch);
  for (int f = threadIdx.x; f < K; f += blockDim.x) Q[f] = Q_s[f];
}

}  // namespace snmfnat

using namespace snmfnat;

extern "C" {

int snmfnat_stft_fft(snmfnat_ctx* ctx, const double* s, int64_t len, int sz, int shift, int fftlen, int DCbin,
                     const double* win, double preemph, double* S_mag, double* S_phase) {
  SN_API_BEGIN
  SN_REQUIRE(ctx && s && win && S_mag && S_phase, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(len >= 0 && sz > 0 && shift > 0 && fftlen >= sz && fftlen % 2 == 0, SNMFNAT_EINVAL, "bad STFT geometry");
  SN_CUDA(cudaSetDevice(ctx->device));
  const int half = fftlen / 2 + 1;
  const int64_t frame_num = len / shift;                                   // stft_fft.m:17
  const int64_t n_valid = (len - fftlen > 1) ? (len - fftlen - 2) / shift + 1 : 0;   // iterations of the loop at :21
  SN_REQUIRE(n_valid <= frame_num, SNMFNAT_EINVAL, "signal too short for its own frame count");
  std::fill(S_mag, S_mag + (size_t)half * frame_num, 0.0);
  std::fill(S_phase, S_phase + (size_t)half * frame_num, 0.0);
  if (n_valid == 0) return SNMFNAT_OK;
  cudaStream_t st = ctx->stream;
  DevBuf<double> ds, dwin, frames, mag, ph;
  DevBuf<double2> Y;
  ds.alloc(len); dwin.alloc(sz); frames.alloc((size_t)n_valid * fftlen); Y.alloc((size_t)n_valid * half);
  mag.alloc((size_t)n_valid * half); ph.alloc((size_t)n_valid * half);
  SN_CUDA(cudaMemcpyAsync(ds.p, s, len * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dwin.p, win, sz * sizeof(double), cudaMemcpyHostToDevice, st));
  const int blocks = (int)std::min<int64_t>(n_valid, (int64_t)ctx->sm_count * 8);
  frame_signal_kernel<<<blocks, 256, 0, st>>>(ds.p, sz, shift, fftlen, preemph, dwin.p, n_valid, frames.p);
  count_launch(ctx);
  FftPlans fft;
  fft.create(ctx, fftlen, n_valid);
  SN_CUFFT(cufftExecD2Z(fft.fwd, frames.p, reinterpret_cast<cufftDoubleComplex*>(Y.p)));
  stft_fft_post_kernel<<<blocks, 256, 0, st>>>(Y.p, half, n_valid, DCbin, mag.p, ph.p);
  count_launch(ctx);
  check_launch(ctx, "stft_fft");
  SN_CUDA(cudaMemcpyAsync(S_mag, mag.p, mag.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaMemcpyAsync(S_phase, ph.p, ph.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  SN_API_END
}

int snmfnat_synth_ifft_buff(snmfnat_ctx* ctx, const double* TF_mag, const double* TF_phase, int freq_num, int frame_num,
                            int sz, int fftlen, const double* win, double preemph, int DCbin_back, double pow_,
                            double* s_buff) {
  SN_API_BEGIN
  SN_REQUIRE(ctx && TF_mag && win && s_buff, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(fftlen % 2 == 0 && sz > 0 && sz <= fftlen && frame_num >= 0 && pow_ > 0, SNMFNAT_EINVAL, "bad geometry");
  const int half = fftlen / 2 + 1;
  SN_REQUIRE(freq_num == half || freq_num == fftlen, SNMFNAT_EINVAL,
             "TF_mag must have fftlen/2+1 or fftlen rows (synth_ifft_buff.m:13-19)");
  SN_REQUIRE(freq_num == fftlen || TF_phase != nullptr, SNMFNAT_EINVAL, "TF_phase is NULL");
  if (frame_num == 0) return SNMFNAT_OK;
  SN_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevBuf<double> dm, dp, dwin, frames;
  DevBuf<double2> Z;
  dm.alloc((size_t)freq_num * frame_num); dwin.alloc(sz); frames.alloc((size_t)frame_num * fftlen);
  Z.alloc((size_t)frame_num * half);
  SN_CUDA(cudaMemcpyAsync(dm.p, TF_mag, dm.n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (freq_num == half) {
    dp.alloc(dm.n);
    SN_CUDA(cudaMemcpyAsync(dp.p, TF_phase, dm.n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  SN_CUDA(cudaMemcpyAsync(dwin.p, win, sz * sizeof(double), cudaMemcpyHostToDevice, st));
  const int blocks = (int)std::min<int64_t>(((int64_t)frame_num * half + 255) / 256, (int64_t)ctx->sm_count * 8);
  synth_pre_kernel<<<blocks, 256, 0, st>>>(dm.p, dp.p, freq_num, fftlen, frame_num, DCbin_back, pow_, Z.p);
  count_launch(ctx);
  FftPlans fft;
  fft.create(ctx, fftlen, frame_num);
  SN_CUFFT(cufftExecZ2D(fft.inv, reinterpret_cast<cufftDoubleComplex*>(Z.p), frames.p));
  StftGeom g{};
  g.sz = sz; g.fftlen = fftlen; g.half = half; g.preemph = preemph;
  launch_synth_window(ctx, g, frames.p, dwin.p, frame_num);
  SN_CUDA(cudaMemcpy2DAsync(s_buff, (size_t)sz * sizeof(double), frames.p, (size_t)fftlen * sizeof(double),
                            (size_t)sz * sizeof(double), frame_num, cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  SN_API_END
}

int snmfnat_blk_sparse(snmfnat_ctx* ctx, const double* X, const double* D, const double* r_blk, int K, int l,
                       const snmfnat_params* p, double* Q, double* r_blk_out) {
  SN_API_BEGIN
  SN_REQUIRE(ctx && X && D && r_blk && p && Q && r_blk_out, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(K > 0 && p->P_len_l > 0 && p->P_len_k >= 2 && p->P_len_k + p->DCbin <= K && p->blk_gap >= 1 &&
                 (p->blk_gap & 1), SNMFNAT_EINVAL, "bad blk_sparse geometry");
  SN_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int P = p->P_len_l;
  OnlineScalars sc{};
  sc.flr = p->nonzerofloor; sc.DCbin = p->DCbin; sc.P_len_k = p->P_len_k; sc.P_len_l = P; sc.blk_gap = p->blk_gap;
  sc.alpha_p = p->alpha_p;
  DevBuf<double> dX, dD, rb, dQ;
  dX.alloc(K); dD.alloc(K); rb.alloc((size_t)K * P); dQ.alloc(K);
  SN_CUDA(cudaMemcpyAsync(dX.p, X, K * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(dD.p, D, K * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(rb.p, r_blk, (size_t)K * P * sizeof(double), cudaMemcpyHostToDevice, st));
  const size_t smem = ((size_t)4 * K + 64) * sizeof(double);
  if (smem > 48 * 1024) SN_CUDA(cudaFuncSetAttribute(blk_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  blk_sparse_kernel<<<1, 256, smem, st>>>(dX.p, dD.p, rb.p, K, l, sc, dQ.p);
  count_launch(ctx);
  check_launch(ctx, "blk_sparse_kernel");
  SN_CUDA(cudaMemcpyAsync(Q, dQ.p, K * sizeof(double), cudaMemcpyDeviceToHost, st));
  // r_blk_out = [r_blk(:,2:P) SNR_local]: slots 1..P-1 then slot 0
  if (P > 1)
    SN_CUDA(cudaMemcpyAsync(r_blk_out, rb.p + K, (size_t)K * (P - 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaMemcpyAsync(r_blk_out + (size_t)K * (P - 1), rb.p, K * sizeof(double), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  SN_API_END
}

}  // extern "C"
