// Per-hop entry (latency path): one audio stream whose struct g (src/init_buff.m) lives on the device; every call of
// snmfnat_stream_step is one call of bnmf_sep_event_RT_IS16 (src/bnmf_sep_event_RT_IS16.m:1-423).
#include <climits>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "state.cuh"

using namespace snmfnat;

struct snmfnat_stream {
  snmfnat_ctx* ctx = nullptr;
  Config cfg;
  SlotBuffers sb;
  DevBuf<double> y, frame, Ym, Yp, Xt, cls, xt_host_dev;
  DevBuf<double> Ysep;   // Mel mode: the separation input of the current frame [LD1]
  DevBuf<double2> Yc;
  FftPlans fft;
  int last_l = 0;
  // DFT mode with a "B_Mel_d" slot that differs from B_DFT_d (bnmf_sep_event_RT_IS16.m:328 [sic]): see snmfnat_stream_create
  bool fix_from_mel = false, fix_synced = false;
};

namespace snmfnat {

// per-class reconstruction with the un-normalised bases (bnmf_sep_event_RT_IS16.m:159-200):
//   out[c][f] = sum_{k in class c} B[f,k] * A[k]
__global__ void class_recon_kernel(const double* __restrict__ B, int LDF, int F, const double* __restrict__ A,
                                   const int* __restrict__ lo, const int* __restrict__ hi, double* __restrict__ out) {
  const int c = blockIdx.y;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = lo[c]; k < hi[c]; ++k) s = fma(B[(size_t)k * LDF + f], A[k], s);
    out[(size_t)c * LDF + f] = s;
  }
}

}  // namespace snmfnat

// chronological (oldest first) <-> ring copies of a history matrix with `cols` time slots of `width` doubles
static void ring_get(snmfnat_ctx* ctx, const double* dev, int ld, int width, int cols, int head, double* host) {
  for (int i = 0; i < cols; ++i) {
    const int slot = (head + i) % cols;
    SN_CUDA(cudaMemcpyAsync(host + (size_t)i * width, dev + (size_t)slot * ld, width * sizeof(double),
                            cudaMemcpyDeviceToHost, ctx->stream));
  }
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
}
static void ring_set(snmfnat_ctx* ctx, double* dev, int ld, int width, int cols, const double* host) {
  SN_CUDA(cudaMemcpy2DAsync(dev, (size_t)ld * sizeof(double), host, (size_t)width * sizeof(double),
                            (size_t)width * sizeof(double), cols, cudaMemcpyHostToDevice, ctx->stream));
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
}

// Ad_blk and lambda_d_blk share one ring position: rotate both so that slot 0 is the oldest column
static void normalize_rings(snmfnat_stream* s) {
  snmfnat_ctx* ctx = s->ctx;
  const OnlineDims& d = s->cfg.d;
  int head = 0;
  SN_CUDA(cudaMemcpy(&head, s->sb.ring_head.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (head == 0) return;
  std::vector<double> a((size_t)d.R_a * d.m_a), b((size_t)d.F * d.m_a);
  ring_get(ctx, s->sb.Ad_blk.p, d.R_a, d.R_a, d.m_a, head, a.data());
  ring_get(ctx, s->sb.lam_blk.p, d.LDF, d.F, d.m_a, head, b.data());
  ring_set(ctx, s->sb.Ad_blk.p, d.R_a, d.R_a, d.m_a, a.data());
  ring_set(ctx, s->sb.lam_blk.p, d.LDF, d.F, d.m_a, b.data());
  const int zero = 0;
  SN_CUDA(cudaMemcpy(s->sb.ring_head.p, &zero, sizeof(int), cudaMemcpyHostToDevice));
}

extern "C" {

int snmfnat_stream_create(snmfnat_ctx* ctx, const snmfnat_params* p, const double* win_stft, const double* win_istft,
                          const double* B_Mel_x, const double* B_Mel_d, int n1, const double* B_DFT_x,
                          const double* B_DFT_d, int n2, const double* Ad_blk_init, const double* A_d_init,
                          snmfnat_stream** out) {
  SN_API_BEGIN
  (void)A_d_init;  // g.A_d is never read by the IS16 frame function (bnmf_sep_event_RT_IS16.m:22,395)
  SN_REQUIRE(ctx && p && win_stft && win_istft && B_DFT_x && B_DFT_d && out, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(p->adapt_train_N == 0 || p->R_a == 0 || Ad_blk_init != nullptr, SNMFNAT_EINVAL,
             "Ad_blk_init is required when adaptation is on (init_buff.m:38 draws it with rand)");
  SN_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<snmfnat_stream> s(new snmfnat_stream());
  s->ctx = ctx;
  make_config(ctx, *p, n2, s->cfg);
  const Config& c = s->cfg;
  const bool mel = c.sc.mel_mode != 0;
  if (mel) {
    // B_sep_mode = 'Mel' (filewise_run_IS16.m:46-51): separation and adaptation on the n1-band Mel dictionaries, gain and
    // block sparsity in the DFT domain (bnmf_sep_event_RT_IS16.m:107-119,165-211,295-319), as in the batch entry
    SN_REQUIRE(B_Mel_x && B_Mel_d && n1 > 0 && n1 <= 256, SNMFNAT_EINVAL, "Mel mode needs B_Mel_x / B_Mel_d (n1 <= 256 bands)");
    SN_REQUIRE(c.p.MelConv != 0, SNMFNAT_EUNSUPPORTED, "Mel mode with MelConv = 0 is not supported");
    SN_REQUIRE(c.p.F_order == n1, SNMFNAT_EINVAL, "B_Mel has %d rows but p.F_order = %d", n1, c.p.F_order);
    SN_REQUIRE(c.p.EVENT_NUM == 1 && c.p.NOISE_NUM == 1, SNMFNAT_EUNSUPPORTED,
               "Mel mode through the per-hop entry supports one event and one noise class");
  } else {
    SN_REQUIRE(n1 == n2, SNMFNAT_EUNSUPPORTED, "DFT mode expects the Mel slots to hold the DFT bases (n1 == n2)");
  }
  s->sb.alloc(1, c.d);
  // B_DFT_d is the adaptable basis; the "B_Mel_d" slot supplies the never-updated columns (:328 [sic])
  s->sb.set_bases(ctx, B_DFT_x, (!mel && B_Mel_d) ? B_Mel_d : B_DFT_d);
  if (mel) {
    std::vector<double> M((size_t)c.d.F * n1);
    mel_matrix_host(c.p.fs, n1, c.p.fftlength, 1.0, c.p.fs / 2.0, M.data());     // init_buff.m:60-62
    s->sb.set_mel(ctx, n1, B_Mel_x, B_Mel_d, M.data());
    s->Ysep.alloc(s->sb.LD1);
  }
  std::vector<int> order(1, 0);
  if (c.sc.adapt_train_N) s->sb.set_ad_init(ctx, Ad_blk_init, 0, 1, order);
  s->sb.win_stft.alloc(c.g.sz); s->sb.win_istft.alloc(c.g.sz);
  SN_CUDA(cudaMemcpy(s->sb.win_stft.p, win_stft, c.g.sz * sizeof(double), cudaMemcpyHostToDevice));
  SN_CUDA(cudaMemcpy(s->sb.win_istft.p, win_istft, c.g.sz * sizeof(double), cudaMemcpyHostToDevice));
  s->sb.reset(ctx);
  if (!mel && B_Mel_d && B_Mel_d != B_DFT_d) {
    // The adaptable atoms (columns < R_a) start from B_DFT_d; the never-updated ones stay B_Mel_d(:, R_a+1:end) in BOTH
    // ping-pong buffers: the reference re-assembles [B_rem, B_new, B_Mel_d(:, R_a+1:end)] on every update (:328,336), so
    // from the first update on the fixed columns are B_Mel_d's whatever B_DFT_d held there.  Before the first update the
    // H-solve reads B_DFT_d as given, hence buffer 0 (the active one) takes all of it.
    upload_basis(ctx, B_DFT_d, c.d.F, c.d.R_d, c.d.LDF, s->sb.Bd0.p);
    s->fix_from_mel = true;
  }
  const int nh = INT_MAX;
  const long long fb = 0;
  SN_CUDA(cudaMemcpy(s->sb.n_hops.p, &nh, sizeof(int), cudaMemcpyHostToDevice));
  SN_CUDA(cudaMemcpy(s->sb.frame_base.p, &fb, sizeof(long long), cudaMemcpyHostToDevice));
  s->y.alloc(c.g.sz); s->frame.alloc(c.g.fftlen); s->Yc.alloc(c.g.half); s->Ym.alloc(c.d.LDF); s->Yp.alloc(c.d.LDF);
  s->Xt.alloc(c.d.LDF);
  s->cls.alloc((size_t)2 * SNMFNAT_MAX_CLASSES * c.d.LDF);
  s->Ym.zero(ctx->stream); s->Yp.zero(ctx->stream); s->Xt.zero(ctx->stream);
  s->fft.create(ctx, c.g.fftlen, 1);
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = s.release();
  SN_API_END
}

int snmfnat_stream_destroy(snmfnat_stream* s) {
  SN_API_BEGIN
  if (s) {
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    delete s;
  }
  SN_API_END
}

// ISTFT of one magnitude vector with the phase of the current frame -> framelength doubles * overlapscale
static void istft_one(snmfnat_stream* s, const double* mag_dev, double* host_out) {
  snmfnat_ctx* ctx = s->ctx;
  const Config& c = s->cfg;
  DevBuf<double2> Z;
  Z.alloc(c.g.half);
  SN_CUDA(cudaMemcpyAsync(Z.p, s->Yc.p, c.g.half * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream));
  launch_istft_pre(ctx, c.g, Z.p, mag_dev, 1);
  SN_CUFFT(cufftExecZ2D(s->fft.inv, reinterpret_cast<cufftDoubleComplex*>(Z.p), s->frame.p));
  launch_synth_window(ctx, c.g, s->frame.p, s->sb.win_istft.p, 1);
  SN_CUDA(cudaMemcpyAsync(host_out, s->frame.p, c.g.sz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < c.g.sz; ++i) host_out[i] *= c.g.overlapscale;   // bnmf_sep_event_RT_IS16.m:354,360,363
}

int snmfnat_stream_step(snmfnat_stream* s, const double* y, int l, const double* h_init, double* x_tilde,
                        double* x_hat_i, double* d_hat_i) {
  SN_API_BEGIN
  SN_REQUIRE(s && y && h_init && x_tilde, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(l >= 1, SNMFNAT_EINVAL, "l is the 1-based hop index");
  snmfnat_ctx* ctx = s->ctx;
  SN_CUDA(cudaSetDevice(ctx->device));
  const Config& c = s->cfg;
  cudaStream_t st = ctx->stream;
  SN_REQUIRE(hsolve_smem_bytes(c.d) <= (size_t)ctx->max_smem_optin, SNMFNAT_EUNSUPPORTED, "basis too large");
  const int loff = 1 - l;  // kernels compute l = g_step + 1 - l_offset with g_step = 0
  SN_CUDA(cudaMemcpyAsync(s->sb.l_offset.p, &loff, sizeof(int), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(s->y.p, y, c.g.sz * sizeof(double), cudaMemcpyHostToDevice, st));
  SN_CUDA(cudaMemcpyAsync(s->sb.h_init.p, h_init, c.d.R * sizeof(double), cudaMemcpyHostToDevice, st));
  // STFT (:66-78)
  launch_frame_one(ctx, c.g, s->y.p, s->sb.win_stft.p, s->frame.p);
  SN_CUFFT(cufftExecD2Z(s->fft.fwd, s->frame.p, reinterpret_cast<cufftDoubleComplex*>(s->Yc.p)));
  launch_stft_post(ctx, c.g, s->Yc.p, 1, s->Ym.p, s->Yp.p);
  // separation, gain, adaptation (:104-347)
  const SlotState sv = s->sb.view();
  FrameArrays fr{s->Ym.p, s->Xt.p};
  if (!c.sc.mel_mode) {
    launch_hsolve(ctx, c.d, c.sc, sv, fr, s->sb.h_init.p, 1, 0);
    launch_gain(ctx, c.d, c.sc, sv, fr, nullptr, 1, 0);
    launch_wsolve(ctx, c.d, c.sc, sv, nullptr, 1, 0);
  } else {
    const SlotState sm = s->sb.view_mel();
    const OnlineDims dm = s->sb.dims_mel();
    FrameArrays frm{s->Ysep.p, nullptr};
    launch_mel_project(ctx, s->sb.melM.p, s->sb.n1, s->sb.LD1, c.d.F, c.d.LDF, s->Ym.p, 1, s->Ysep.p);
    launch_hsolve(ctx, dm, c.sc, sm, frm, s->sb.h_init.p, 1, 0);
    launch_mel_post(ctx, c.d, sv, s->sb.melM.p, s->sb.n1, s->sb.LD1, s->sb.XhatM.p, s->sb.DhatM.p, s->Ysep.p, 1, 0);
    launch_gain(ctx, c.d, c.sc, sv, fr, nullptr, 1, 0);
    if (c.sc.adapt_train_N) {
      launch_mel_hist(ctx, c.d, sv, s->sb.melM.p, s->sb.n1, s->sb.LD1, s->sb.lam_blk_mel.p, 1, 0);
      launch_wsolve(ctx, dm, c.sc, sm, nullptr, 1, 0);
    }
  }
  // ISTFT (:349-363)
  istft_one(s, s->Xt.p, x_tilde);
  if ((x_hat_i || d_hat_i) && c.sc.mel_mode) {
    // one event / one noise class: the class reconstructions are melmat' * (B_Mel * A) = what mel_post_kernel left in
    // Xhat / Dhat (bnmf_sep_event_RT_IS16.m:165-202)
    if (x_hat_i) istft_one(s, s->sb.Xhat.p, x_hat_i);
    if (d_hat_i) istft_one(s, s->sb.Dhat.p, d_hat_i);
  } else if (x_hat_i || d_hat_i) {
    const snmfnat_params& p = c.p;
    std::vector<int> lo, hi;
    for (int i = 0; i < p.EVENT_NUM; ++i) {   // :159-164
      lo.push_back(p.EVENT_RANK[i] - 1);
      hi.push_back(i == p.EVENT_NUM - 1 ? p.R_x : p.EVENT_RANK[i + 1] - 1);
    }
    for (int i = 0; i < p.NOISE_NUM; ++i) {   // :180-185
      lo.push_back(p.R_x + p.NOISE_RANK[i] - 1);
      hi.push_back(i == p.NOISE_NUM - 1 ? p.R_x + p.R_d : p.R_x + p.NOISE_RANK[i + 1] - 1);
    }
    const int nc = (int)lo.size();
    DevBuf<int> dlo, dhi;
    DevBuf<double> Bfull;
    dlo.alloc(nc); dhi.alloc(nc);
    SN_CUDA(cudaMemcpyAsync(dlo.p, lo.data(), nc * sizeof(int), cudaMemcpyHostToDevice, st));
    SN_CUDA(cudaMemcpyAsync(dhi.p, hi.data(), nc * sizeof(int), cudaMemcpyHostToDevice, st));
    // [B_x B_d] with the noise basis as it was during this hop's separation: the W-solve may just have flipped the
    // buffers, the pre-update copy is then in the other one
    int sel = 0, upd = 0;
    SN_CUDA(cudaMemcpyAsync(&sel, s->sb.bd_sel.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SN_CUDA(cudaMemcpyAsync(&upd, s->sb.do_update.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SN_CUDA(cudaStreamSynchronize(st));
    const int used = upd ? (sel ^ 1) : sel;
    Bfull.alloc((size_t)c.d.R * c.d.LDF);
    SN_CUDA(cudaMemcpyAsync(Bfull.p, s->sb.Bx.p, (size_t)c.d.R_x * c.d.LDF * sizeof(double), cudaMemcpyDeviceToDevice, st));
    SN_CUDA(cudaMemcpyAsync(Bfull.p + (size_t)c.d.R_x * c.d.LDF, used ? s->sb.Bd1.p : s->sb.Bd0.p,
                            (size_t)c.d.R_d * c.d.LDF * sizeof(double), cudaMemcpyDeviceToDevice, st));
    class_recon_kernel<<<dim3(4, nc), 256, 0, st>>>(Bfull.p, c.d.LDF, c.d.F, s->sb.A.p, dlo.p, dhi.p, s->cls.p);
    count_launch(ctx);
    check_launch(ctx, "class_recon_kernel");
    for (int i = 0; i < p.EVENT_NUM; ++i)
      if (x_hat_i) istft_one(s, s->cls.p + (size_t)i * c.d.LDF, x_hat_i + (size_t)i * c.g.sz);
    for (int i = 0; i < p.NOISE_NUM; ++i)
      if (d_hat_i) istft_one(s, s->cls.p + (size_t)(p.EVENT_NUM + i) * c.d.LDF, d_hat_i + (size_t)i * c.g.sz);
  }
  int flag = 0, sel = 0;
  SN_CUDA(cudaMemcpyAsync(&flag, s->sb.err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (s->fix_from_mel && !s->fix_synced) SN_CUDA(cudaMemcpyAsync(&sel, s->sb.bd_sel.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SN_CUDA(cudaStreamSynchronize(st));
  if (sel == 1) {
    // first update done: from now on the never-updated columns are B_Mel_d's in both buffers (the W-solve only rewrites
    // columns < R_a, so buffer 0 still carries B_DFT_d's fixed columns until they are replaced here, before its next use)
    const size_t off = (size_t)c.d.R_a * c.d.LDF, cnt = (size_t)(c.d.R_d - c.d.R_a) * c.d.LDF;
    SN_CUDA(cudaMemcpyAsync(s->sb.Bd0.p + off, s->sb.Bd_fix.p + off, cnt * sizeof(double), cudaMemcpyDeviceToDevice, st));
    s->fix_synced = true;
  }
  SN_REQUIRE(flag == 0, SNMFNAT_ENUMERIC,
             "an all-zero activation row was selected for adaptation (bnmf_sep_event_RT_IS16.m:292 vs :323)");
  s->last_l = l;
  SN_API_END
}

static const double* current_bd(snmfnat_stream* s) {
  int sel = 0;
  SN_CUDA(cudaMemcpy(&sel, s->sb.bd_sel.p, sizeof(int), cudaMemcpyDeviceToHost));
  return sel ? s->sb.Bd1.p : s->sb.Bd0.p;
}

int snmfnat_stream_get(snmfnat_stream* s, const char* field, double* buf, int64_t n) {
  SN_API_BEGIN
  SN_REQUIRE(s && field && buf, SNMFNAT_EINVAL, "NULL argument");
  snmfnat_ctx* ctx = s->ctx;
  SN_CUDA(cudaSetDevice(ctx->device));
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
  const Config& c = s->cfg;
  const OnlineDims& d = c.d;
  const std::string f(field);
  auto need = [&](int64_t k) { SN_REQUIRE(n >= k, SNMFNAT_EINVAL, "buffer too small for '%s': need %lld doubles", field, (long long)k); };
  auto vec = [&](const double* dev, int len) {
    need(len);
    SN_CUDA(cudaMemcpy(buf, dev, len * sizeof(double), cudaMemcpyDeviceToHost));
  };
  auto ival = [&](const int* dev) {
    need(1);
    int v = 0;
    SN_CUDA(cudaMemcpy(&v, dev, sizeof(int), cudaMemcpyDeviceToHost));
    buf[0] = v;
  };
  const bool mel = c.sc.mel_mode != 0;
  const int n1 = s->sb.n1, LD1 = s->sb.LD1;
  if (mel && f == "B_Mel_d") {          // the adapted dictionary of Mel mode (n1 x R_d)
    need((int64_t)n1 * d.R_d);
    int sel = 0;
    SN_CUDA(cudaMemcpy(&sel, s->sb.bd_sel.p, sizeof(int), cudaMemcpyDeviceToHost));
    download_basis(ctx, sel ? s->sb.BdM1.p : s->sb.BdM0.p, n1, d.R_d, LD1, buf);
  } else if (mel && f == "B_Mel_x") {
    need((int64_t)n1 * d.R_x);
    download_basis(ctx, s->sb.BxM.p, n1, d.R_x, LD1, buf);
  }
  else if (f == "B_DFT_d") { need((int64_t)d.F * d.R_d); download_basis(ctx, mel ? s->sb.Bd_fix.p : current_bd(s), d.F, d.R_d, d.LDF, buf); }
  else if (f == "B_Mel_d") { need((int64_t)d.F * d.R_d); download_basis(ctx, s->sb.Bd_fix.p, d.F, d.R_d, d.LDF, buf); }
  else if (f == "B_DFT_x" || f == "B_Mel_x") { need((int64_t)d.F * d.R_x); download_basis(ctx, s->sb.Bx.p, d.F, d.R_x, d.LDF, buf); }
  else if (f == "Ad_blk") {
    need((int64_t)d.R_a * d.m_a);
    int head = 0;
    SN_CUDA(cudaMemcpy(&head, s->sb.ring_head.p, sizeof(int), cudaMemcpyDeviceToHost));
    ring_get(ctx, s->sb.Ad_blk.p, d.R_a, d.R_a, d.m_a, head, buf);
  } else if (f == "lambda_d_blk") {
    need((int64_t)d.F * d.m_a);
    int head = 0;
    SN_CUDA(cudaMemcpy(&head, s->sb.ring_head.p, sizeof(int), cudaMemcpyDeviceToHost));
    ring_get(ctx, s->sb.lam_blk.p, d.LDF, d.F, d.m_a, head, buf);
  } else if (f == "r_blk") {
    need((int64_t)d.F * d.P_len_l);
    int pos = 0;
    SN_CUDA(cudaMemcpy(&pos, s->sb.rblk_pos.p, sizeof(int), cudaMemcpyDeviceToHost));
    ring_get(ctx, s->sb.r_blk.p, d.LDF, d.F, d.P_len_l, pos, buf);
  }
  else if (f == "lambda_dav") vec(s->sb.lambda_dav.p, d.F);
  else if (f == "Xm_tilde") vec(s->sb.Xm_tilde_prev.p, d.F);
  else if (f == "Ym") vec(s->Ym.p, d.F);
  else if (f == "Yp") vec(s->Yp.p, d.F);
  else if (f == "A") vec(s->sb.A.p, d.R);
  else if (f == "Q") vec(s->sb.Q.p, d.F);
  else if (f == "G") vec(s->sb.G.p, d.F);
  else if (f == "Xm_hat") vec(s->sb.Xhat.p, d.F);
  else if (f == "Dm_hat") vec(s->sb.Dhat.p, d.F);
  else if (f == "update_switch") ival(s->sb.update_switch.p);
  else if (f == "stats") {
    need(5);
    int hi = 0, ga = 0, nu = 0, wi = 0;
    SN_CUDA(cudaMemcpy(&hi, s->sb.h_iters.p, sizeof(int), cudaMemcpyDeviceToHost));
    SN_CUDA(cudaMemcpy(&ga, s->sb.gated.p, sizeof(int), cudaMemcpyDeviceToHost));
    SN_CUDA(cudaMemcpy(&nu, s->sb.n_up.p, sizeof(int), cudaMemcpyDeviceToHost));
    SN_CUDA(cudaMemcpy(&wi, s->sb.w_iters.p, sizeof(int), cudaMemcpyDeviceToHost));
    buf[0] = hi; buf[1] = ga; buf[2] = nu; buf[3] = wi;
    SN_CUDA(cudaMemcpy(&buf[4], s->sb.h_cost.p, sizeof(double), cudaMemcpyDeviceToHost));
  }
  else fail(SNMFNAT_EINVAL, "unknown field '%s'", field);
  SN_API_END
}

int snmfnat_stream_set(snmfnat_stream* s, const char* field, const double* buf, int64_t n) {
  SN_API_BEGIN
  SN_REQUIRE(s && field && buf, SNMFNAT_EINVAL, "NULL argument");
  snmfnat_ctx* ctx = s->ctx;
  SN_CUDA(cudaSetDevice(ctx->device));
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
  const OnlineDims& d = s->cfg.d;
  const std::string f(field);
  auto need = [&](int64_t k) { SN_REQUIRE(n >= k, SNMFNAT_EINVAL, "buffer too small for '%s': need %lld doubles", field, (long long)k); };
  const int zero = 0;
  if (f == "B_DFT_d") {   // e.g. the carry-over of B_D_u.mat (src/NTF_sep_event_RT.m:28-38)
    need((int64_t)d.F * d.R_d);
    upload_basis(ctx, buf, d.F, d.R_d, d.LDF, s->sb.Bd0.p);
    upload_basis(ctx, buf, d.F, d.R_d, d.LDF, s->sb.Bd1.p);
    SN_CUDA(cudaMemcpy(s->sb.bd_sel.p, &zero, sizeof(int), cudaMemcpyHostToDevice));
  } else if (f == "B_Mel_d") {
    need((int64_t)d.F * d.R_d);
    upload_basis(ctx, buf, d.F, d.R_d, d.LDF, s->sb.Bd_fix.p);
  } else if (f == "Ad_blk") {
    need((int64_t)d.R_a * d.m_a);
    normalize_rings(s);
    ring_set(ctx, s->sb.Ad_blk.p, d.R_a, d.R_a, d.m_a, buf);
  } else if (f == "lambda_d_blk") {
    need((int64_t)d.F * d.m_a);
    normalize_rings(s);
    ring_set(ctx, s->sb.lam_blk.p, d.LDF, d.F, d.m_a, buf);
  } else if (f == "r_blk") {
    need((int64_t)d.F * d.P_len_l);
    ring_set(ctx, s->sb.r_blk.p, d.LDF, d.F, d.P_len_l, buf);
    SN_CUDA(cudaMemcpy(s->sb.rblk_pos.p, &zero, sizeof(int), cudaMemcpyHostToDevice));
  } else if (f == "lambda_dav") {
    need(d.F);
    SN_CUDA(cudaMemcpy(s->sb.lambda_dav.p, buf, d.F * sizeof(double), cudaMemcpyHostToDevice));
  } else if (f == "Xm_tilde") {
    need(d.F);
    SN_CUDA(cudaMemcpy(s->sb.Xm_tilde_prev.p, buf, d.F * sizeof(double), cudaMemcpyHostToDevice));
  } else if (f == "update_switch") {
    need(1);
    const int v = (int)buf[0];
    SN_CUDA(cudaMemcpy(s->sb.update_switch.p, &v, sizeof(int), cudaMemcpyHostToDevice));
  } else fail(SNMFNAT_EINVAL, "field '%s' cannot be set", field);
  SN_API_END
}

}  // extern "C"
