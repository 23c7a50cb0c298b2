// STFT / ISTFT kernels around cuFFT (float64).  All of these are HBM-bound streaming kernels: one pass over
// the frame arrays with 16-byte vector accesses on the wide side.
#include "stft.cuh"

namespace snmfnat {

// ---------------------------------------------------------------------------------------------------
// framing: the 640-sample queue of filewise_run_IS16.m:121-122 at hop l holds samples
// [l*shift - sz, l*shift); flush hops (l > floor(len/shift)) are all-zero frames (:111-113).
// Pre-emphasis is per frame with zero state (bnmf_sep_event_RT_IS16.m:67), then the analysis window (:68)
// and zero padding to fftlen (:69).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
frame_pcm_kernel(StftGeom g, UttTables ut, const int16_t* __restrict__ pcm, const double* __restrict__ win,
                 double* __restrict__ frames) {
  const int u = blockIdx.y;
  const int nh = ut.n_hops[u];
  const long long len = ut.len[u];
  const long long n_full = len / g.shift;
  const int16_t* __restrict__ x = pcm + ut.pcm_off[u];
  const int pairs = g.fftlen / 2;
  for (int l = blockIdx.x + 1; l <= nh; l += gridDim.x) {
    double2* __restrict__ out = reinterpret_cast<double2*>(frames + (size_t)(ut.frame_base[u] + l - 1) * g.fftlen);
    const bool live = (l <= n_full);
    const long long start = (long long)l * g.shift - g.sz;
    for (int pi = threadIdx.x; pi < pairs; pi += blockDim.x) {
      double2 o = make_double2(0.0, 0.0);
      if (live) {
        const int i = 2 * pi;
        double v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ii = i + e;
          double val = 0.0;
          if (ii < g.sz) {
            const long long pos = start + ii;
            const double cur = (pos >= 0) ? (double)x[pos] : 0.0;
            const double prev = (ii > 0 && pos - 1 >= 0) ? (double)x[pos - 1] : 0.0;
            val = win[ii] * (cur - g.preemph * prev);
          }
          v[e] = val;
        }
        o = make_double2(v[0], v[1]);
      }
      out[pi] = o;
    }
  }
}

void launch_frame_pcm(snmfnat_ctx* ctx, const StftGeom& g, const UttTables& ut, const int16_t* pcm, const double* win,
                      double* frames) {
  if (ut.n_utt <= 0) return;
  SN_REQUIRE(g.fftlen % 2 == 0, SNMFNAT_EINVAL, "fftlength must be even");
  const int gx = ut.max_hops < 64 ? (ut.max_hops > 0 ? ut.max_hops : 1) : 64;
  for (int u0 = 0; u0 < ut.n_utt; u0 += 65535) {
    UttTables t = ut;
    const int nu = (ut.n_utt - u0 < 65535) ? ut.n_utt - u0 : 65535;
    t.pcm_off += u0; t.len += u0; t.frame_base += u0; t.n_hops += u0; t.out_off += u0;
    frame_pcm_kernel<<<dim3(gx, nu), 256, 0, ctx->stream>>>(g, t, pcm, win, frames);
    count_launch(ctx);
  }
  check_launch(ctx, "frame_pcm_kernel");
}

__global__ void frame_one_kernel(StftGeom g, const double* __restrict__ y, const double* __restrict__ win,
                                 double* __restrict__ frame) {
  for (int i = threadIdx.x; i < g.fftlen; i += blockDim.x) {
    double v = 0.0;
    if (i < g.sz) v = win[i] * (y[i] - g.preemph * (i > 0 ? y[i - 1] : 0.0));
    frame[i] = v;
  }
}
void launch_frame_one(snmfnat_ctx* ctx, const StftGeom& g, const double* y, const double* win, double* frame) {
  frame_one_kernel<<<1, 256, 0, ctx->stream>>>(g, y, win, frame);
  count_launch(ctx);
  check_launch(ctx, "frame_one_kernel");
}

// ---------------------------------------------------------------------------------------------------
// Ym = abs(Y).^pow ; Ym(1:DCbin) = 0 ; Ym += floor     (bnmf_sep_event_RT_IS16.m:71-78)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mag_pow(double2 c, double pw) {
  const double p2 = fma(c.x, c.x, c.y * c.y);
  if (pw == 2.0) return p2;
  if (pw == 1.0) return sqrt(p2);
  return pow(sqrt(p2), pw);
}

__global__ void __launch_bounds__(256)
stft_post_kernel(StftGeom g, const double2* __restrict__ Y, long long nf, double* __restrict__ Ym,
                 double* __restrict__ Yp) {
  const int ppf = g.LDF / 2;  // output pairs per frame
  const long long total = nf * ppf;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long fr = idx / ppf;
    const int f = 2 * (int)(idx - fr * ppf);
    const double2* __restrict__ yrow = Y + (size_t)fr * g.half;
    double o[2], ph[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ff = f + e;
      double val = 0.0, p = 0.0;
      if (ff < g.half) {
        const double2 c = yrow[ff];
        val = (ff < g.DCbin ? 0.0 : mag_pow(c, g.pow_)) + g.flr;
        if (Yp) p = atan2(c.y, c.x);
      }
      o[e] = val;
      ph[e] = p;
    }
    *reinterpret_cast<double2*>(Ym + (size_t)fr * g.LDF + f) = make_double2(o[0], o[1]);
    if (Yp) *reinterpret_cast<double2*>(Yp + (size_t)fr * g.LDF + f) = make_double2(ph[0], ph[1]);
  }
}

static int grid_for(snmfnat_ctx* ctx, long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  const long long cap = (long long)ctx->sm_count * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

void launch_stft_post(snmfnat_ctx* ctx, const StftGeom& g, const double2* Y, long long nf, double* Ym, double* Yp) {
  if (nf <= 0) return;
  stft_post_kernel<<<grid_for(ctx, nf * (g.LDF / 2), 256), 256, 0, ctx->stream>>>(g, Y, nf, Ym, Yp);
  count_launch(ctx);
  check_launch(ctx, "stft_post_kernel");
}

// ---------------------------------------------------------------------------------------------------
// TF_mag(1:DCbin_back)=0 ; TF_mag.^(1/pow) ; TF = mag .* exp(1i*phase)       (synth_ifft_buff.m:10-18)
// exp(1i*angle(Y)) is evaluated as Y/|Y| (angle(0) = 0 -> 1).  The Hermitian mirror (:16-17) is implied by C2R.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
istft_pre_kernel(StftGeom g, double2* __restrict__ Y, const double* __restrict__ Xt, long long nf) {
  const long long total = nf * g.half;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long fr = idx / g.half;
    const int f = (int)(idx - fr * g.half);
    const double2 c = Y[idx];
    double mag = 0.0;
    if (f >= g.DCbin_back) {
      const double x = Xt[(size_t)fr * g.LDF + f];
      mag = (g.pow_ == 2.0) ? sqrt(x) : ((g.pow_ == 1.0) ? x : pow(x, 1.0 / g.pow_));
    }
    const double a = sqrt(fma(c.x, c.x, c.y * c.y));
    double2 z;
    if (a > 0.0) z = make_double2(mag * (c.x / a), mag * (c.y / a));
    else z = make_double2(mag, 0.0);
    Y[idx] = z;
  }
}
void launch_istft_pre(snmfnat_ctx* ctx, const StftGeom& g, double2* Y, const double* Xt, long long nf) {
  if (nf <= 0) return;
  istft_pre_kernel<<<grid_for(ctx, nf * g.half, 256), 256, 0, ctx->stream>>>(g, Y, Xt, nf);
  count_launch(ctx);
  check_launch(ctx, "istft_pre_kernel");
}

// ---------------------------------------------------------------------------------------------------
// s_proc = real(ifft(TF))(1:sz) .* win ; de-emphasis          (synth_ifft_buff.m:20-26)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
synth_window_kernel(StftGeom g, double* __restrict__ frames, const double* __restrict__ win, long long nf) {
  const double inv = 1.0 / g.fftlen;
  for (long long fr = blockIdx.x; fr < nf; fr += gridDim.x) {
    double* __restrict__ row = frames + (size_t)fr * g.fftlen;
    for (int i = threadIdx.x; i < g.sz; i += blockDim.x) row[i] = row[i] * inv * win[i];
    if (g.preemph != 0.0) {
      __syncthreads();
      if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int i = 0; i < g.sz; ++i) {
          acc = row[i] + g.preemph * acc;
          row[i] = acc;
        }
      }
      __syncthreads();
    }
  }
}
void launch_synth_window(snmfnat_ctx* ctx, const StftGeom& g, double* frames, const double* win, long long nf) {
  if (nf <= 0) return;
  long long b = nf < (long long)ctx->sm_count * 8 ? nf : (long long)ctx->sm_count * 8;
  synth_window_kernel<<<(int)b, 256, 0, ctx->stream>>>(g, frames, win, nf);
  count_launch(ctx);
  check_launch(ctx, "synth_window_kernel");
}

// ---------------------------------------------------------------------------------------------------
// overlap-add + int16 (filewise_run_IS16.m:146,162-165): output block j (l = j+delay+1) is the first `shift`
// samples of the OLA buffer after adding frame l; only frames with l > delay were ever added.
// `windowed` != 0: frames already hold s_proc (synth_window_kernel ran); else apply /fftlen and the window here.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int16_t to_int16_sat(double v) {
  if (!(v == v)) return 0;
  double r = round(v);  // half away from zero, like MATLAB's double->int16
  if (r > 32767.0) r = 32767.0;
  if (r < -32768.0) r = -32768.0;
  return (int16_t)r;
}

__global__ void __launch_bounds__(256)
ola_int16_kernel(StftGeom g, UttTables ut, const double* __restrict__ frames, const double* __restrict__ win,
                 int windowed, int16_t* __restrict__ out) {
  const int u = blockIdx.y;
  const int nh = ut.n_hops[u];
  const long long nout = (long long)(nh - g.delay) * g.shift;
  int16_t* __restrict__ o = out + ut.out_off[u];
  const double* __restrict__ fb = frames + (size_t)ut.frame_base[u] * g.fftlen;
  const int nov = (g.sz + g.shift - 1) / g.shift;
  const double inv = 1.0 / g.fftlen;
  for (long long s_out = blockIdx.x * (long long)blockDim.x + threadIdx.x; s_out < nout;
       s_out += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(s_out / g.shift);
    const int s = (int)(s_out - (long long)j * g.shift);
    const int l = j + g.delay + 1;
    double acc = 0.0;
    for (int i = nov - 1; i >= 0; --i) {  // oldest frame first, like the running buffer
      const int lf = l - i;
      const int idx = i * g.shift + s;
      if (lf > g.delay && lf >= 1 && idx < g.sz) {
        double x = fb[(size_t)(lf - 1) * g.fftlen + idx];
        if (!windowed) x = x * inv * win[idx];
        acc += x * g.overlapscale;
      }
    }
    o[s_out] = to_int16_sat(acc);
  }
}

void launch_ola_int16(snmfnat_ctx* ctx, const StftGeom& g, const UttTables& ut, const double* frames,
                        const double* win, int windowed, int16_t* out) {
  if (ut.n_utt <= 0) return;
  long long per = (long long)ut.max_hops * g.shift;
  int gx = (int)((per + 255) / 256);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  for (int u0 = 0; u0 < ut.n_utt; u0 += 65535) {
    UttTables t = ut;
    const int nu = (ut.n_utt - u0 < 65535) ? ut.n_utt - u0 : 65535;
    t.pcm_off += u0; t.len += u0; t.frame_base += u0; t.n_hops += u0; t.out_off += u0;
    ola_int16_kernel<<<dim3(gx, nu), 256, 0, ctx->stream>>>(g, t, frames, win, windowed, out);
    count_launch(ctx);
  }
  check_launch(ctx, "ola_int16_kernel");
}

// ---------------------------------------------------------------------------------------------------
void FftPlans::create(snmfnat_ctx* ctx, int n, long long frames) {
  destroy();
  SN_REQUIRE(frames > 0 && frames < (1ll << 40), SNMFNAT_EINVAL, "bad frame count %lld", frames);
  // 64-bit plans: frames * n passes 2^31 elements at about 2 700 CHiME-length utterances in one batch
  long long nn[1] = {n};
  size_t ws_f = 0, ws_i = 0;
  SN_CUFFT(cufftCreate(&fwd));
  cufftResult rf = cufftMakePlanMany64(fwd, 1, nn, nullptr, 1, n, nullptr, 1, n / 2 + 1, CUFFT_D2Z, frames, &ws_f);
  if (rf != CUFFT_SUCCESS) {
    cufftDestroy(fwd);
    fwd = 0;
    fail(rf == CUFFT_ALLOC_FAILED ? SNMFNAT_ENOMEM : SNMFNAT_EUNSUPPORTED,
         "cuFFT cannot plan %lld transforms of length %d in one batch (cufft error %d): split the corpus into smaller batches",
         frames, n, (int)rf);
  }
  SN_CUFFT(cufftCreate(&inv));
  cufftResult ri = cufftMakePlanMany64(inv, 1, nn, nullptr, 1, n / 2 + 1, nullptr, 1, n, CUFFT_Z2D, frames, &ws_i);
  if (ri != CUFFT_SUCCESS) {
    cufftDestroy(fwd);
    cufftDestroy(inv);
    fwd = inv = 0;
    fail(ri == CUFFT_ALLOC_FAILED ? SNMFNAT_ENOMEM : SNMFNAT_EUNSUPPORTED,
         "cuFFT cannot plan %lld inverse transforms of length %d in one batch (cufft error %d): split the corpus into smaller batches",
         frames, n, (int)ri);
  }
  SN_CUFFT(cufftSetStream(fwd, ctx->stream));
  SN_CUFFT(cufftSetStream(inv, ctx->stream));
  nf = frames;
  fftlen = n;
  ok = true;
}
void FftPlans::destroy() {
  if (ok) {
    cufftDestroy(fwd);
    cufftDestroy(inv);
  }
  ok = false;
  fwd = inv = 0;
}

}  // namespace snmfnat
