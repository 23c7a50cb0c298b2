#include <cmath>
#include <cstring>
#include "state.cuh"

namespace snmfnat {

void make_config(snmfnat_ctx* ctx, const snmfnat_params& p, int n2, Config& cfg) {
  (void)ctx;
  // Configurations the reference's IS16 frame function cannot run itself (index errors):
  SN_REQUIRE(p.blk_len_sep == 1 && p.blk_hop_sep == 1 && p.Splice == 0, SNMFNAT_EUNSUPPORTED,
             "blk_len_sep>1 / Splice>0 are not runnable in bnmf_sep_event_RT_IS16 (Ym is local, :86-100)");
  SN_REQUIRE(p.fftlength >= p.framelength && p.fftlength % 2 == 0 && p.framelength > 0 && p.frameshift > 0,
             SNMFNAT_EINVAL, "bad frame geometry");
  SN_REQUIRE(n2 == p.fftlength / 2 + 1, SNMFNAT_EINVAL, "basis has %d rows, expected fftlength/2+1 = %d", n2,
             p.fftlength / 2 + 1);
  SN_REQUIRE(p.B_sep_mode == SNMFNAT_SEP_DFT || p.B_sep_mode == SNMFNAT_SEP_MEL, SNMFNAT_EINVAL, "bad B_sep_mode");
  SN_REQUIRE(p.B_sep_mode == SNMFNAT_SEP_DFT || p.MelConv == 1, SNMFNAT_EUNSUPPORTED,
             "B_sep_mode='Mel' is implemented for MelConv=1 (Mel -> DFT conversion of the separated spectra, :165-172)");
  SN_REQUIRE(p.cf == SNMFNAT_CF_KL, SNMFNAT_EUNSUPPORTED, "online path implements cf='kl' only");
  SN_REQUIRE((p.basis_update_N == 0 && p.basis_update_E == 0) || p.B_sep_mode == SNMFNAT_SEP_DFT, SNMFNAT_EUNSUPPORTED,
             "basis_update_N/E (W-update inside the separation solve) is implemented for B_sep_mode='DFT'");
  SN_REQUIRE(p.R_x > 0 && p.R_d > 0 && p.R_a >= 0 && p.R_a <= p.R_d, SNMFNAT_EINVAL, "bad ranks");
  SN_REQUIRE(p.m_a > 0 && p.P_len_l > 0 && p.max_iter >= 0, SNMFNAT_EINVAL, "bad m_a / P_len_l / max_iter");
  SN_REQUIRE(p.DCbin >= 0 && p.DCbin < n2 && p.DCbin_back >= 0 && p.DCbin_back <= n2, SNMFNAT_EINVAL, "bad DCbin");
  SN_REQUIRE(p.EVENT_NUM >= 1 && p.EVENT_NUM <= SNMFNAT_MAX_CLASSES && p.NOISE_NUM >= 1 &&
                 p.NOISE_NUM <= SNMFNAT_MAX_CLASSES, SNMFNAT_EINVAL, "bad class counts");
  SN_REQUIRE(p.pow > 0, SNMFNAT_EINVAL, "pow must be positive");
  // class c owns the atoms RANK(c) .. RANK(c+1)-1 (bnmf_sep_event_RT_IS16.m:159-164,181-186): the reference indexes out of
  // range when a class starts beyond the dictionary (e.g. initial_setting_SNMF_Techwin_201603_RT.m: EVENT_RANK = [1 21 41]
  // with R_x = 20), and atoms in front of the first class would silently drop out of the reconstruction
  SN_REQUIRE(p.EVENT_RANK[0] == 1 && p.NOISE_RANK[0] == 1, SNMFNAT_EUNSUPPORTED,
             "EVENT_RANK(1) / NOISE_RANK(1) must be 1 (atoms in front of the first class are not supported)");
  for (int i = 1; i < p.EVENT_NUM; ++i)
    SN_REQUIRE(p.EVENT_RANK[i] > p.EVENT_RANK[i - 1] && p.EVENT_RANK[i] <= p.R_x, SNMFNAT_EINVAL,
               "EVENT_RANK(%d) = %d: classes must increase and start within R_x = %d", i + 1, p.EVENT_RANK[i], p.R_x);
  for (int i = 1; i < p.NOISE_NUM; ++i)
    SN_REQUIRE(p.NOISE_RANK[i] > p.NOISE_RANK[i - 1] && p.NOISE_RANK[i] <= p.R_d, SNMFNAT_EINVAL,
               "NOISE_RANK(%d) = %d: classes must increase and start within R_d = %d", i + 1, p.NOISE_RANK[i], p.R_d);
  cfg.p = p;
  OnlineDims& d = cfg.d;
  d.F = n2;
  d.LDF = pad_ld(n2);
  d.R_x = p.R_x; d.R_d = p.R_d; d.R = p.R_x + p.R_d; d.R_a = p.R_a; d.m_a = p.m_a; d.P_len_l = p.P_len_l;
  // bnmf_sep_event_RT_IS16.m:125-139: an if / elseif chain, so basis_update_N wins when both are set
  if (p.basis_update_N) { d.upd0 = p.R_x; d.upd1 = p.R_x + p.R_d; }
  else if (p.basis_update_E) { d.upd0 = 0; d.upd1 = p.R_x; }
  OnlineScalars& s = cfg.sc;
  s.flr = p.nonzerofloor;
  s.sparsity = p.sparsity; s.conv_eps = p.conv_eps; s.max_iter = p.max_iter; s.cost_check = p.cost_check;
  s.DCbin = p.DCbin; s.init_N_len = p.init_N_len; s.adapt_train_N = p.adapt_train_N && p.R_a > 0;
  s.blk_sparse = p.blk_sparse; s.P_len_k = p.P_len_k; s.P_len_l = p.P_len_l; s.blk_gap = p.blk_gap;
  s.alpha_p = p.alpha_p; s.alpha_eta = p.alpha_eta; s.alpha_d = p.alpha_d; s.beta = p.beta; s.beta_max = p.beta_max;
  s.Ar_up = p.Ar_up; s.enhance_method = p.ENHANCE_METHOD;
  s.update_period = (int)std::floor(p.overlap_m_a * (double)p.m_a);  // bnmf_sep_event_RT_IS16.m:293
  s.mel_mode = p.B_sep_mode == SNMFNAT_SEP_MEL;
  StftGeom& g = cfg.g;
  g.sz = p.framelength; g.shift = p.frameshift; g.fftlen = p.fftlength; g.half = n2; g.LDF = d.LDF; g.delay = p.delay;
  g.preemph = p.preemph; g.pow_ = p.pow; g.flr = p.nonzerofloor; g.overlapscale = p.overlapscale;
  g.DCbin = p.DCbin; g.DCbin_back = p.DCbin_back;
}

void upload_basis(snmfnat_ctx* ctx, const double* host, int F, int R, int LDF, double* dev) {
  SN_CUDA(cudaMemsetAsync(dev, 0, (size_t)R * LDF * sizeof(double), ctx->stream));
  SN_CUDA(cudaMemcpy2DAsync(dev, (size_t)LDF * sizeof(double), host, (size_t)F * sizeof(double),
                            (size_t)F * sizeof(double), R, cudaMemcpyHostToDevice, ctx->stream));
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
}
void download_basis(snmfnat_ctx* ctx, const double* dev, int F, int R, int LDF, double* host) {
  SN_CUDA(cudaMemcpy2DAsync(host, (size_t)F * sizeof(double), dev, (size_t)LDF * sizeof(double),
                            (size_t)F * sizeof(double), R, cudaMemcpyDeviceToHost, ctx->stream));
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
}

void SlotBuffers::alloc(int S_, const OnlineDims& d_) {
  S = S_;
  d = d_;
  const size_t LDF = d.LDF;
  Bx.alloc((size_t)d.R_x * LDF);
  Bd_fix.alloc((size_t)d.R_d * LDF);
  Bd0.alloc((size_t)S * d.R_d * LDF);
  Bd1.alloc((size_t)S * d.R_d * LDF);
  Ad_blk.alloc((size_t)S * d.m_a * d.R_a);
  lam_blk.alloc((size_t)S * d.m_a * LDF);
  r_blk.alloc((size_t)S * d.P_len_l * LDF);
  lambda_dav.alloc((size_t)S * LDF);
  Xm_tilde_prev.alloc((size_t)S * LDF);
  A.alloc((size_t)S * d.R);
  Xhat.alloc((size_t)S * LDF);
  Dhat.alloc((size_t)S * LDF);
  Q.alloc((size_t)S * LDF);
  G.alloc((size_t)S * LDF);
  h_cost.alloc(S);
  h_init.alloc(d.R);
  rblk_pos.alloc(S); bd_sel.alloc(S); ring_head.alloc(S); update_switch.alloc(S); h_iters.alloc(S); gated.alloc(S);
  do_update.alloc(S); n_up.alloc(S); w_iters.alloc(S); err_flag.alloc(1);
  idx_up.alloc((size_t)S * (d.R_a > 0 ? d.R_a : 1));
  idx_rem.alloc((size_t)S * (d.R_a > 0 ? d.R_a : 1));
  l_offset.alloc(S); n_hops.alloc(S); frame_base.alloc(S);
  stats.alloc(8);
  if (d.upd1 > d.upd0) semi_w.alloc((size_t)S * (d.upd1 - d.upd0) * LDF);
}

void SlotBuffers::set_bases(snmfnat_ctx* ctx, const double* B_x, const double* B_d) {
  upload_basis(ctx, B_x, d.F, d.R_x, d.LDF, Bx.p);
  upload_basis(ctx, B_d, d.F, d.R_d, d.LDF, Bd_fix.p);
  if (hsolve_ms_supported(ctx, d)) {
    ms_colstat.alloc(2 * 152);
    ms_perm.alloc((size_t)16 * S);
    ms_perm_step.alloc(16);
    ms_ticket.alloc(16);
    ws_perm.alloc((size_t)16 * S);
    ws_perm_step.alloc(16);
    w_last.alloc(S);
    launch_ms_colstat(ctx, d, Bx.p, Bd_fix.p, ms_colstat.p);
    SN_CUDA(cudaStreamSynchronize(ctx->stream));
  }
}

void SlotBuffers::set_mel(snmfnat_ctx* ctx, int n1_, const double* B_Mel_x, const double* B_Mel_d, const double* melmat) {
  n1 = n1_;
  LD1 = pad_ld(n1);
  const size_t L1 = LD1;
  melM.alloc((size_t)n1 * d.LDF);
  BxM.alloc((size_t)d.R_x * L1);
  BdM_fix.alloc((size_t)d.R_d * L1);
  BdM0.alloc((size_t)S * d.R_d * L1);
  BdM1.alloc((size_t)S * d.R_d * L1);
  lam_blk_mel.alloc((size_t)S * d.m_a * L1);
  XhatM.alloc((size_t)S * L1);
  DhatM.alloc((size_t)S * L1);
  upload_basis(ctx, melmat, d.F, n1, d.LDF, melM.p);      // column b of melmat (n2 long) -> row b of M
  upload_basis(ctx, B_Mel_x, n1, d.R_x, LD1, BxM.p);
  upload_basis(ctx, B_Mel_d, n1, d.R_d, LD1, BdM_fix.p);
}

OnlineDims SlotBuffers::dims_mel() const {
  OnlineDims m = d;
  m.F = n1;
  m.LDF = LD1;
  return m;
}

SlotState SlotBuffers::view_mel() const {
  SlotState v = view();
  v.Bx = BxM.p;
  v.Bd_fix = BdM_fix.p;
  v.Bd[0] = BdM0.p; v.Bd[1] = BdM1.p;
  v.lam_blk = lam_blk_mel.p;
  v.Xhat = XhatM.p; v.Dhat = DhatM.p;
  return v;
}

void SlotBuffers::set_ad_init(snmfnat_ctx* ctx, const double* Ad_blk_init, int64_t stride, int n_utt,
                              const std::vector<int>& first_utt) {
  // MATLAB R_a x m_a column-major == [m_a][R_a] time-slot major: the ring layout, oldest column first
  const size_t per = (size_t)d.m_a * d.R_a;
  if (per == 0) return;
  n_utt_ad = n_utt;
  Ad_init.alloc((size_t)n_utt * per);
  if (stride == (int64_t)per) {
    SN_CUDA(cudaMemcpyAsync(Ad_init.p, Ad_blk_init, (size_t)n_utt * per * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    for (int u = 0; u < n_utt; ++u) {
      const double* src = Ad_blk_init + (stride ? (size_t)u * stride : 0);
      SN_CUDA(cudaMemcpyAsync(Ad_init.p + (size_t)u * per, src, per * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  first_utt_dev.alloc(S);
  SN_CUDA(cudaMemcpyAsync(first_utt_dev.p, first_utt.data(), (size_t)S * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
}

__global__ void fill_int_kernel(int* p, int n, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}
// Ad_blk of slot s <- Ad_init of utterance first_utt[s]
__global__ void gather_ad_kernel(const double* __restrict__ ad_init, const int* __restrict__ first_utt, double* __restrict__ ad_blk,
                                 size_t per, int S) {
  const size_t total = per * (size_t)S;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t s = i / per;
    ad_blk[i] = ad_init[(size_t)first_utt[s] * per + (i - s * per)];
  }
}
// One block per event: re-run init_buff (src/init_buff.m:17-62) for the slot, keeping Bd / bd_sel.
__global__ void chain_boundary_kernel(const int* __restrict__ ev, OnlineDims d, double* lam_blk_mel, int LD1, double* lam_blk, double* r_blk, double* lambda_dav,
                                      double* Xm_tilde_prev, double* ad_blk, const double* __restrict__ ad_init, int* rblk_pos,
                                      int* ring_head, int* update_switch, int* l_offset, int* n_hops) {
  const int slot = ev[4 * blockIdx.x], utt = ev[4 * blockIdx.x + 1], step0 = ev[4 * blockIdx.x + 2], nh = ev[4 * blockIdx.x + 3];
  const size_t LDF = d.LDF;
  double* p = lam_blk + (size_t)slot * d.m_a * LDF;
  for (size_t i = threadIdx.x; i < (size_t)d.m_a * LDF; i += blockDim.x) p[i] = 0.0;
  if (lam_blk_mel) {
    p = lam_blk_mel + (size_t)slot * d.m_a * LD1;
    for (size_t i = threadIdx.x; i < (size_t)d.m_a * LD1; i += blockDim.x) p[i] = 0.0;
  }
  p = r_blk + (size_t)slot * d.P_len_l * LDF;
  for (size_t i = threadIdx.x; i < (size_t)d.P_len_l * LDF; i += blockDim.x) p[i] = 0.0;
  for (size_t i = threadIdx.x; i < LDF; i += blockDim.x) {
    lambda_dav[(size_t)slot * LDF + i] = 0.0;
    Xm_tilde_prev[(size_t)slot * LDF + i] = 0.0;
  }
  const size_t per = (size_t)d.m_a * d.R_a;
  for (size_t i = threadIdx.x; i < per; i += blockDim.x) ad_blk[(size_t)slot * per + i] = ad_init[(size_t)utt * per + i];
  if (threadIdx.x == 0) {
    rblk_pos[slot] = 0;
    ring_head[slot] = 0;
    update_switch[slot] = 1;   // init_buff.m:41
    l_offset[slot] = step0;    // l restarts at 1
    n_hops[slot] = nh;
  }
}

void SlotBuffers::chain_boundary(snmfnat_ctx* ctx, const int* events_dev, int n_events) {
  if (n_events <= 0) return;
  chain_boundary_kernel<<<n_events, 512, 0, ctx->stream>>>(events_dev, d, mel() ? lam_blk_mel.p : nullptr, LD1, lam_blk.p, r_blk.p, lambda_dav.p, Xm_tilde_prev.p,
                                                            Ad_blk.p, Ad_init.p, rblk_pos.p, ring_head.p, update_switch.p,
                                                            l_offset.p, n_hops.p);
  count_launch(ctx);
  check_launch(ctx, "chain_boundary_kernel");
}

__global__ void broadcast_kernel(const double* __restrict__ src, double* __restrict__ dst, size_t per, int S) {
  const size_t total = per * (size_t)S;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i % per];
}

void SlotBuffers::reset(snmfnat_ctx* ctx) {
  cudaStream_t st = ctx->stream;
  lam_blk.zero(st); r_blk.zero(st); lambda_dav.zero(st); Xm_tilde_prev.zero(st);
  A.zero(st); Xhat.zero(st); Dhat.zero(st); Q.zero(st); G.zero(st); h_cost.zero(st);
  rblk_pos.zero(st); bd_sel.zero(st); ring_head.zero(st); h_iters.zero(st); gated.zero(st); do_update.zero(st); n_up.zero(st);
  w_iters.zero(st); err_flag.zero(st); idx_up.zero(st); idx_rem.zero(st); stats.zero(st);
  if (ms_perm.n) {
    ms_ticket.zero(st);
    SN_CUDA(cudaMemsetAsync(ms_perm_step.p, 0xff, 16 * sizeof(int), st));   // -1: no order computed yet
    SN_CUDA(cudaMemsetAsync(ws_perm_step.p, 0xff, 16 * sizeof(int), st));
    w_last.zero(st);
  }
  fill_int_kernel<<<(S + 255) / 256, 256, 0, st>>>(update_switch.p, S, 1);  // init_buff.m:41
  count_launch(ctx);
  const size_t per = (size_t)d.R_d * d.LDF;
  int blocks = (int)((per * S + 255) / 256);
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  broadcast_kernel<<<blocks, 256, 0, st>>>(Bd_fix.p, Bd0.p, per, S);
  broadcast_kernel<<<blocks, 256, 0, st>>>(Bd_fix.p, Bd1.p, per, S);  // columns >= R_a must be valid in both buffers
  count_launch(ctx, 2);
  if (mel()) {
    const size_t perm = (size_t)d.R_d * LD1;
    int bm = (int)((perm * S + 255) / 256);
    if (bm > ctx->sm_count * 16) bm = ctx->sm_count * 16;
    broadcast_kernel<<<bm, 256, 0, st>>>(BdM_fix.p, BdM0.p, perm, S);
    broadcast_kernel<<<bm, 256, 0, st>>>(BdM_fix.p, BdM1.p, perm, S);
    count_launch(ctx, 2);
    lam_blk_mel.zero(st); XhatM.zero(st); DhatM.zero(st);
  }
  if (Ad_blk.n && Ad_init.n) {
    gather_ad_kernel<<<blocks, 256, 0, st>>>(Ad_init.p, first_utt_dev.p, Ad_blk.p, (size_t)d.m_a * d.R_a, S);
    count_launch(ctx);
  }
  check_launch(ctx, "state reset");
}

SlotState SlotBuffers::view() const {
  SlotState v{};
  v.S = S;
  v.Bx = Bx.p;
  v.Bd_fix = Bd_fix.p;
  v.bdfix_stride = 0;
  v.Bd[0] = Bd0.p; v.Bd[1] = Bd1.p;
  v.bd_sel = bd_sel.p;
  v.Ad_blk = Ad_blk.p; v.lam_blk = lam_blk.p; v.ring_head = ring_head.p; v.r_blk = r_blk.p; v.rblk_pos = rblk_pos.p;
  v.lambda_dav = lambda_dav.p; v.Xm_tilde_prev = Xm_tilde_prev.p; v.update_switch = update_switch.p;
  v.A = A.p; v.Xhat = Xhat.p; v.Dhat = Dhat.p; v.Q = Q.p; v.G = G.p;
  v.h_iters = h_iters.p; v.h_cost = h_cost.p; v.gated = gated.p; v.do_update = do_update.p; v.n_up = n_up.p;
  v.idx_up = idx_up.p; v.idx_rem = idx_rem.p; v.w_iters = w_iters.p; v.err_flag = err_flag.p;
  v.l_offset = l_offset.p; v.n_hops = n_hops.p; v.frame_base = frame_base.p;
  v.stats = stats.p;
  v.ms_colstat = ms_colstat.p;
  v.ms_perm = ms_perm.p; v.ms_perm_step = ms_perm_step.p; v.ms_ticket = ms_ticket.p; v.ms_perm_stride = S;
  v.ws_perm = ws_perm.p; v.ws_perm_step = ws_perm_step.p; v.w_last = w_last.p;
  v.semi_w = semi_w.p;
  return v;
}

}  // namespace snmfnat
