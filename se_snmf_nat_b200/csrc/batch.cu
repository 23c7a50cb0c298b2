// Whole-batch entry: the hop loops of filewise_run_IS16.m:86-169 for n_utt utterances advanced in lock step on
// one GPU.  STFT of every frame up front (batched cuFFT), then one {H-solve, gain, W-solve} kernel triple per
// hop over all still-active utterances, then ISTFT + overlap-add of every frame.
#include <algorithm>
#include <array>
#include <memory>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include "state.cuh"

using namespace snmfnat;

struct snmfnat_batch {
  snmfnat_ctx* ctx = nullptr;
  Config cfg;
  int n_utt = 0;
  std::vector<int64_t> len, out_len, out_off_h, pcm_off_h;
  std::vector<int> n_hops_u;       // per utterance
  std::vector<int> order;          // slot -> first utterance of the slot's unit (units sorted by hops, longest first)
  std::vector<int> slot_of;        // utterance -> slot
  int n_slots = 0;                 // units: single utterances, or chains of utterances (B_D_u.mat carry-over)
  std::vector<int> events;         // chain boundaries {slot, utt, step0, n_hops}, sorted by step0
  std::vector<int> ev_begin;       // ev_begin[g] .. ev_begin[g+1]: events of global step g
  DevBuf<int> d_events, d_loff0, d_nhops0;
  std::vector<long long> frame_base_u;
  std::vector<int> active_at;      // active_at[g] = number of active slots at global step g
  int max_hops = 0;
  long long NF = 0;                // total frames
  int64_t pcm_total = 0, out_total = 0;
  SlotBuffers sb;
  DevBuf<int16_t> pcm, out;
  DevBuf<long long> d_pcm_off, d_len, d_frame_base_u, d_out_off;
  DevBuf<int> d_n_hops_u;
  DevBuf<double> frames, Ym, Xt;
  DevBuf<double> Ysep;             // Mel mode: the separation input of every frame [NF][LD1]
  DevBuf<double2> Yc;
  DevBuf<double> trA, trQ, trG;
  DevBuf<int> trInfo;
  bool trace = false, uploaded = false, ran = false;
  FftPlans fft;
  int64_t launches_last = 0;
  // profiling: events around every launch of the hop loop
  bool profile = false;
  std::vector<cudaEvent_t> ev;
  double prof_ms[6] = {0, 0, 0, 0, 0, 0};
  int64_t prof_cnt[6] = {0, 0, 0, 0, 0, 0};
  // interleaved slot groups on their own CUDA streams (slot s belongs to group s % n_groups)
  int n_groups = 3;                // snmfnat_batch_create picks it from the number of slots (see there)
  std::vector<cudaStream_t> gstream;
  std::vector<cudaEvent_t> gdone;
  cudaEvent_t fork_ev = nullptr;
  ~snmfnat_batch() {
    for (auto e : ev) cudaEventDestroy(e);
    for (auto e : gdone) cudaEventDestroy(e);
    for (auto st : gstream) cudaStreamDestroy(st);
    if (fork_ev) cudaEventDestroy(fork_ev);
  }
};

static UttTables utt_tables(const snmfnat_batch* b) {
  UttTables t{};
  t.pcm_off = b->d_pcm_off.p; t.len = b->d_len.p; t.frame_base = b->d_frame_base_u.p; t.n_hops = b->d_n_hops_u.p;
  t.out_off = b->d_out_off.p; t.n_utt = b->n_utt; t.max_hops = b->max_hops;
  return t;
}

extern "C" {

int snmfnat_batch_create(snmfnat_ctx* ctx, const snmfnat_params* p, const double* win_stft, const double* win_istft,
                         const double* B_x, const double* B_d, int n2, int n_utt, const int64_t* len,
                         const int32_t* chain_id, const double* h_init, const double* Ad_blk_init, int64_t ad_stride,
                         snmfnat_batch** out) {
  SN_API_BEGIN
  SN_REQUIRE(ctx && p && win_stft && win_istft && B_x && B_d && len && h_init && out, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(n_utt > 0, SNMFNAT_EINVAL, "n_utt must be positive");
  SN_REQUIRE(p->adapt_train_N == 0 || p->R_a == 0 || Ad_blk_init != nullptr, SNMFNAT_EINVAL,
             "Ad_blk_init is required when adaptation is on (init_buff.m:38 draws it with rand)");
  SN_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<snmfnat_batch> b(new snmfnat_batch());
  b->ctx = ctx;
  make_config(ctx, *p, n2, b->cfg);
  const Config& c = b->cfg;
  SN_REQUIRE(hsolve_smem_bytes(c.d) <= (size_t)ctx->max_smem_optin, SNMFNAT_EUNSUPPORTED,
             "basis [B_x B_d] (%d x %d) does not fit the 4-CTA shared-memory H-solve", c.d.F, c.d.R);
  b->n_utt = n_utt;
  b->n_groups = 0;   // chosen below from the number of slots unless SNMFNAT_GROUPS / snmfnat_batch_set_groups says otherwise
  if (const char* e = getenv("SNMFNAT_GROUPS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 16) b->n_groups = v;
  }
  b->len.assign(len, len + n_utt);
  b->n_hops_u.resize(n_utt); b->out_len.resize(n_utt); b->pcm_off_h.resize(n_utt); b->out_off_h.resize(n_utt);
  b->frame_base_u.resize(n_utt);
  int64_t po = 0, oo = 0;
  for (int u = 0; u < n_utt; ++u) {
    SN_REQUIRE(len[u] >= 0 && len[u] / c.g.shift < (1 << 30), SNMFNAT_EINVAL, "bad length of utterance %d", u);
    const int nh = (int)(len[u] / c.g.shift) + c.g.delay + 1;  // filewise_run_IS16.m:102-123
    b->n_hops_u[u] = nh;
    b->out_len[u] = (int64_t)(nh - c.g.delay) * c.g.shift;     // :146,165
    b->pcm_off_h[u] = po;
    po += (len[u] + 7) / 8 * 8;
    b->out_off_h[u] = oo;
    oo += (b->out_len[u] + 7) / 8 * 8;
  }
  b->pcm_total = po; b->out_total = oo;
  // slots: one per unit (an utterance, or a chain of utterances that hand their adapted noise basis on through
  // B_D_u.mat, src/NTF_sep_event_RT.m:28-38,136-139), longest first, so the active set at any step is a prefix
  std::vector<std::vector<int>> units;
  {
    std::vector<std::pair<int, int>> seen;  // chain id -> unit
    for (int u = 0; u < n_utt; ++u) {
      const int cid = chain_id ? chain_id[u] : -1;
      int at = -1;
      if (cid >= 0)
        for (auto& pr : seen)
          if (pr.first == cid) at = pr.second;
      if (at < 0) {
        at = (int)units.size();
        units.emplace_back();
        if (cid >= 0) seen.emplace_back(cid, at);
      }
      units[at].push_back(u);
    }
  }
  const int n_slots = (int)units.size();
  b->n_slots = n_slots;
  // Interleaved slot groups on separate CUDA streams: with ~1000 slots every launch fills the GPU for several waves and 3
  // groups are enough to overlap the H-solve of one group with the W-solves of another; a few hundred slots (one GPU's
  // share of a corpus split over 8) leave every launch with one or two partly filled waves, and more, smaller groups pack
  // them better (measured on one B200: 128 slots 1 196 -> 1 356 xRT, 256 slots 1 500 -> 1 566 xRT with 8 instead of 3 groups;
  // 1024 slots: 1 686 vs 1 677).
  if (b->n_groups == 0) b->n_groups = n_slots <= 320 ? 8 : (n_slots <= 640 ? 6 : 3);
  std::vector<long long> unit_hops(n_slots, 0);
  for (int j = 0; j < n_slots; ++j)
    for (int u : units[j]) unit_hops[j] += b->n_hops_u[u];
  SN_REQUIRE(*std::max_element(unit_hops.begin(), unit_hops.end()) < (1ll << 30), SNMFNAT_EINVAL, "a chain is too long");
  std::vector<int> uorder(n_slots);
  std::iota(uorder.begin(), uorder.end(), 0);
  std::stable_sort(uorder.begin(), uorder.end(), [&](int a, int q) { return unit_hops[a] > unit_hops[q]; });
  b->order.resize(n_slots);
  b->slot_of.resize(n_utt);
  long long fb = 0;
  std::vector<long long> fb_slot(n_slots);
  std::vector<int> nh_slot(n_slots), loff(n_slots, 0), total_slot(n_slots);
  std::vector<std::array<int, 4>> evs;
  for (int s = 0; s < n_slots; ++s) {
    const std::vector<int>& mem = units[uorder[s]];
    b->order[s] = mem[0];
    fb_slot[s] = fb;
    nh_slot[s] = b->n_hops_u[mem[0]];
    total_slot[s] = (int)unit_hops[uorder[s]];
    int step = 0;
    for (size_t j = 0; j < mem.size(); ++j) {
      const int u = mem[j];
      b->slot_of[u] = s;
      b->frame_base_u[u] = fb;          // the frames of a chain are consecutive: frame = frame_base[slot] + step
      if (j > 0) evs.push_back({s, u, step, b->n_hops_u[u]});
      fb += b->n_hops_u[u];
      step += b->n_hops_u[u];
    }
  }
  b->NF = fb;
  b->max_hops = total_slot[0];
  b->active_at.assign(b->max_hops, 0);
  for (int g = 0, s = n_slots; g < b->max_hops; ++g) {
    while (s > 0 && total_slot[s - 1] <= g) --s;
    b->active_at[g] = s;
  }
  std::stable_sort(evs.begin(), evs.end(), [](const std::array<int, 4>& a, const std::array<int, 4>& q) { return a[2] < q[2]; });
  b->ev_begin.assign(b->max_hops + 1, 0);
  for (auto& e : evs) {
    b->events.insert(b->events.end(), e.begin(), e.end());
    b->ev_begin[e[2] + 1]++;
  }
  for (int g = 0; g < b->max_hops; ++g) b->ev_begin[g + 1] += b->ev_begin[g];
  // device buffers
  b->sb.alloc(n_slots, c.d);
  b->sb.set_bases(ctx, B_x, B_d);
  if (c.sc.adapt_train_N) b->sb.set_ad_init(ctx, Ad_blk_init, ad_stride, n_utt, b->order);
  b->sb.win_stft.alloc(c.g.sz); b->sb.win_istft.alloc(c.g.sz);
  SN_CUDA(cudaMemcpy(b->sb.win_stft.p, win_stft, c.g.sz * sizeof(double), cudaMemcpyHostToDevice));
  SN_CUDA(cudaMemcpy(b->sb.win_istft.p, win_istft, c.g.sz * sizeof(double), cudaMemcpyHostToDevice));
  SN_CUDA(cudaMemcpy(b->sb.h_init.p, h_init, c.d.R * sizeof(double), cudaMemcpyHostToDevice));
  b->d_loff0.alloc(n_slots); b->d_nhops0.alloc(n_slots);
  SN_CUDA(cudaMemcpy(b->d_loff0.p, loff.data(), n_slots * sizeof(int), cudaMemcpyHostToDevice));
  SN_CUDA(cudaMemcpy(b->d_nhops0.p, nh_slot.data(), n_slots * sizeof(int), cudaMemcpyHostToDevice));
  SN_CUDA(cudaMemcpy(b->sb.frame_base.p, fb_slot.data(), n_slots * sizeof(long long), cudaMemcpyHostToDevice));
  if (!b->events.empty()) {
    b->d_events.alloc(b->events.size());
    SN_CUDA(cudaMemcpy(b->d_events.p, b->events.data(), b->events.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  b->pcm.alloc((size_t)std::max<int64_t>(po, 8));
  b->out.alloc((size_t)std::max<int64_t>(oo, 8));
  b->d_pcm_off.alloc(n_utt); b->d_len.alloc(n_utt); b->d_frame_base_u.alloc(n_utt); b->d_out_off.alloc(n_utt);
  b->d_n_hops_u.alloc(n_utt);
  {
    std::vector<long long> t(n_utt);
    for (int u = 0; u < n_utt; ++u) t[u] = b->pcm_off_h[u];
    SN_CUDA(cudaMemcpy(b->d_pcm_off.p, t.data(), n_utt * sizeof(long long), cudaMemcpyHostToDevice));
    for (int u = 0; u < n_utt; ++u) t[u] = b->len[u];
    SN_CUDA(cudaMemcpy(b->d_len.p, t.data(), n_utt * sizeof(long long), cudaMemcpyHostToDevice));
    for (int u = 0; u < n_utt; ++u) t[u] = b->out_off_h[u];
    SN_CUDA(cudaMemcpy(b->d_out_off.p, t.data(), n_utt * sizeof(long long), cudaMemcpyHostToDevice));
    SN_CUDA(cudaMemcpy(b->d_frame_base_u.p, b->frame_base_u.data(), n_utt * sizeof(long long), cudaMemcpyHostToDevice));
    SN_CUDA(cudaMemcpy(b->d_n_hops_u.p, b->n_hops_u.data(), n_utt * sizeof(int), cudaMemcpyHostToDevice));
  }
  b->frames.alloc((size_t)b->NF * c.g.fftlen);
  b->Yc.alloc((size_t)b->NF * c.g.half);
  b->Ym.alloc((size_t)b->NF * c.d.LDF);
  b->Xt.alloc((size_t)b->NF * c.d.LDF);
  b->fft.create(ctx, c.g.fftlen, b->NF);
  SN_CUDA(cudaMemset(b->pcm.p, 0, b->pcm.n * sizeof(int16_t)));
  *out = b.release();
  SN_API_END
}

int snmfnat_batch_destroy(snmfnat_batch* b) {
  SN_API_BEGIN
  if (b) {
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    delete b;
  }
  SN_API_END
}

int snmfnat_batch_upload(snmfnat_batch* b, const int16_t* const* pcm) {
  SN_API_BEGIN
  SN_REQUIRE(b && pcm, SNMFNAT_EINVAL, "NULL argument");
  SN_CUDA(cudaSetDevice(b->ctx->device));
  for (int u = 0; u < b->n_utt; ++u)
    if (b->len[u] > 0)
      SN_CUDA(cudaMemcpyAsync(b->pcm.p + b->pcm_off_h[u], pcm[u], b->len[u] * sizeof(int16_t), cudaMemcpyHostToDevice,
                              b->ctx->stream));
  SN_CUDA(cudaStreamSynchronize(b->ctx->stream));
  b->uploaded = true;
  SN_API_END
}

int snmfnat_batch_upload_packed(snmfnat_batch* b, const int16_t* pcm_packed) {
  SN_API_BEGIN
  SN_REQUIRE(b && pcm_packed, SNMFNAT_EINVAL, "NULL argument");
  SN_CUDA(cudaSetDevice(b->ctx->device));
  int64_t off = 0;
  for (int u = 0; u < b->n_utt; ++u) {
    if (b->len[u] > 0)
      SN_CUDA(cudaMemcpyAsync(b->pcm.p + b->pcm_off_h[u], pcm_packed + off, b->len[u] * sizeof(int16_t),
                              cudaMemcpyHostToDevice, b->ctx->stream));
    off += b->len[u];
  }
  b->uploaded = true;
  SN_API_END
}

int snmfnat_batch_enable_trace(snmfnat_batch* b, int on) {
  SN_API_BEGIN
  SN_REQUIRE(b, SNMFNAT_EINVAL, "NULL argument");
  if (on && !b->trace) {
    b->trA.alloc((size_t)b->NF * b->cfg.d.R);
    b->trQ.alloc((size_t)b->NF * b->cfg.d.LDF);
    b->trG.alloc((size_t)b->NF * b->cfg.d.LDF);
    b->trInfo.alloc((size_t)b->NF * 4);
  }
  b->trace = on != 0;
  SN_API_END
}

int snmfnat_batch_run(snmfnat_batch* b) {
  SN_API_BEGIN
  SN_REQUIRE(b, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(b->uploaded, SNMFNAT_EINVAL, "snmfnat_batch_upload must be called before snmfnat_batch_run");
  snmfnat_ctx* ctx = b->ctx;
  SN_CUDA(cudaSetDevice(ctx->device));
  const Config& c = b->cfg;
  const int64_t l0 = ctx->launches;
  const bool prof = b->profile;
  size_t evi = 0;
  if (prof) {
    const size_t need = 3 * (size_t)b->max_hops + 8;
    while (b->ev.size() < need) {
      cudaEvent_t e;
      SN_CUDA(cudaEventCreate(&e));
      b->ev.push_back(e);
    }
  }
  auto mark = [&]() {
    if (prof) SN_CUDA(cudaEventRecord(b->ev[evi++], ctx->stream));
  };
  NvtxRange nvtx_run("snmfnat_batch_run");
  mark();  // 0: start
  b->sb.reset(ctx);
  SN_CUDA(cudaMemcpyAsync(b->sb.l_offset.p, b->d_loff0.p, b->n_slots * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  SN_CUDA(cudaMemcpyAsync(b->sb.n_hops.p, b->d_nhops0.p, b->n_slots * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  const UttTables ut = utt_tables(b);
  // STFT of every frame
  nvtxRangePushA("stft");
  launch_frame_pcm(ctx, c.g, ut, b->pcm.p, b->sb.win_stft.p, b->frames.p);
  SN_CUFFT(cufftExecD2Z(b->fft.fwd, b->frames.p, reinterpret_cast<cufftDoubleComplex*>(b->Yc.p)));
  launch_stft_post(ctx, c.g, b->Yc.p, b->NF, b->Ym.p, nullptr);
  const bool mel = c.sc.mel_mode != 0;
  if (mel) {
    SN_REQUIRE(b->sb.mel(), SNMFNAT_EINVAL, "B_sep_mode='Mel': call snmfnat_batch_set_mel before snmfnat_batch_run");
    launch_mel_project(ctx, b->sb.melM.p, b->sb.n1, b->sb.LD1, c.d.F, c.d.LDF, b->Ym.p, b->NF, b->Ysep.p);
  }
  nvtxRangePop();
  mark();  // 1: STFT done
  nvtxRangePushA("hop loop: hsolve / gain / wsolve");
  // hop loop
  const SlotState st = b->sb.view();
  FrameArrays fr{b->Ym.p, b->Xt.p};
  TraceArrays tr{b->trA.p, b->trQ.p, b->trG.p, b->trInfo.p};
  const TraceArrays* trp = b->trace ? &tr : nullptr;
  SlotState st_mel{};
  FrameArrays fr_mel{nullptr, nullptr};
  if (mel) {
    st_mel = b->sb.view_mel();
    fr_mel = FrameArrays{b->Ysep.p, nullptr};
  }
  const int NG = prof ? 1 : b->n_groups;   // the per-class event profile needs the launches serialised on one stream
  if (NG > 1) {
    while ((int)b->gstream.size() < NG) {
      cudaStream_t q;
      SN_CUDA(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
      b->gstream.push_back(q);
      cudaEvent_t e;
      SN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      b->gdone.push_back(e);
    }
    if (!b->fork_ev) SN_CUDA(cudaEventCreateWithFlags(&b->fork_ev, cudaEventDisableTiming));
    SN_CUDA(cudaEventRecord(b->fork_ev, ctx->stream));
    for (int q = 0; q < NG; ++q) SN_CUDA(cudaStreamWaitEvent(b->gstream[q], b->fork_ev, 0));
  }
  // the launch helpers read ctx->stream: point it at the group's stream inside the loop and restore it on EVERY exit path
  // (a throwing launch must not leave the shared context on a stream that snmfnat_batch_destroy is about to destroy)
  struct StreamGuard {
    snmfnat_ctx* c;
    cudaStream_t saved;
    ~StreamGuard() { c->stream = saved; }
  } stream_guard{ctx, ctx->stream};
  cudaStream_t main_stream = ctx->stream;
  for (int g = 0; g < b->max_hops; ++g) {
    const int na = b->active_at[g];
    for (int q = 0; q < NG; ++q) {
      const int nq = na > q ? (na - q + NG - 1) / NG : 0;
      if (NG > 1) ctx->stream = b->gstream[q];
      // slots whose chain moves on to its next file at this step
      for (int e = b->ev_begin[g]; e < b->ev_begin[g + 1]; ++e)
        if (b->events[4 * (size_t)e] % NG == q) b->sb.chain_boundary(ctx, b->d_events.p + 4 * (size_t)e, 1);
      if (nq == 0) continue;
      OnlineDims dq = c.d;
      dq.slot0 = q;
      dq.slot_stride = NG;
      if (!mel) {
        launch_hsolve(ctx, dq, c.sc, st, fr, b->sb.h_init.p, nq, g);
        mark();
        const int na_next = g + 1 < b->max_hops ? b->active_at[g + 1] : 0;
        launch_gain(ctx, dq, c.sc, st, fr, trp, nq, g, na_next > q ? (na_next - q + NG - 1) / NG : 0);
        mark();
        launch_wsolve(ctx, dq, c.sc, st, trp, nq, g);
        mark();
      } else {
        // the two solves run on the Mel-sized view of the state (n1 bands); gain, block sparsity and the gate stay in
        // the DFT domain (bnmf_sep_event_RT_IS16.m:107-119,165-211,295-319)
        OnlineDims dm = b->sb.dims_mel();
        dm.slot0 = q;
        dm.slot_stride = NG;
        launch_hsolve(ctx, dm, c.sc, st_mel, fr_mel, b->sb.h_init.p, nq, g);
        mark();
        launch_mel_post(ctx, dq, st, b->sb.melM.p, b->sb.n1, b->sb.LD1, b->sb.XhatM.p, b->sb.DhatM.p, b->Ysep.p, nq, g);
        launch_gain(ctx, dq, c.sc, st, fr, trp, nq, g);
        mark();
        if (c.sc.adapt_train_N) {
          launch_mel_hist(ctx, dq, st, b->sb.melM.p, b->sb.n1, b->sb.LD1, b->sb.lam_blk_mel.p, nq, g);
          launch_wsolve(ctx, dm, c.sc, st_mel, trp, nq, g);
        }
        mark();
      }
    }
  }
  ctx->stream = main_stream;
  if (NG > 1)
    for (int q = 0; q < NG; ++q) {
      SN_CUDA(cudaEventRecord(b->gdone[q], b->gstream[q]));
      SN_CUDA(cudaStreamWaitEvent(main_stream, b->gdone[q], 0));
    }
  nvtxRangePop();
  // ISTFT + overlap-add
  NvtxRange nvtx_istft("istft + overlap-add");
  launch_istft_pre(ctx, c.g, b->Yc.p, b->Xt.p, b->NF);
  SN_CUFFT(cufftExecZ2D(b->fft.inv, reinterpret_cast<cufftDoubleComplex*>(b->Yc.p), b->frames.p));
  int windowed = 0;
  if (c.g.preemph != 0.0) {
    launch_synth_window(ctx, c.g, b->frames.p, b->sb.win_istft.p, b->NF);
    windowed = 1;
  }
  launch_ola_int16(ctx, c.g, ut, b->frames.p, b->sb.win_istft.p, windowed, b->out.p);
  mark();  // end
  b->launches_last = ctx->launches - l0;
  b->ran = true;
  SN_API_END
}

static void check_err_flag(snmfnat_batch* b) {
  int flag = 0;
  SN_CUDA(cudaMemcpyAsync(&flag, b->sb.err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, b->ctx->stream));
  SN_CUDA(cudaStreamSynchronize(b->ctx->stream));
  SN_REQUIRE(flag == 0, SNMFNAT_ENUMERIC,
             "an all-zero activation row was selected for adaptation (reference dimension mismatch, "
             "bnmf_sep_event_RT_IS16.m:292 vs :323)");
}

int snmfnat_batch_download(snmfnat_batch* b, int16_t* const* out) {
  SN_API_BEGIN
  SN_REQUIRE(b && out, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(b->ran, SNMFNAT_EINVAL, "snmfnat_batch_run has not been called");
  SN_CUDA(cudaSetDevice(b->ctx->device));
  for (int u = 0; u < b->n_utt; ++u)
    SN_CUDA(cudaMemcpyAsync(out[u], b->out.p + b->out_off_h[u], b->out_len[u] * sizeof(int16_t), cudaMemcpyDeviceToHost,
                            b->ctx->stream));
  check_err_flag(b);
  SN_API_END
}

int snmfnat_batch_download_packed(snmfnat_batch* b, int16_t* out_packed) {
  SN_API_BEGIN
  SN_REQUIRE(b && out_packed, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(b->ran, SNMFNAT_EINVAL, "snmfnat_batch_run has not been called");
  SN_CUDA(cudaSetDevice(b->ctx->device));
  int64_t off = 0;
  for (int u = 0; u < b->n_utt; ++u) {
    SN_CUDA(cudaMemcpyAsync(out_packed + off, b->out.p + b->out_off_h[u], b->out_len[u] * sizeof(int16_t),
                            cudaMemcpyDeviceToHost, b->ctx->stream));
    off += b->out_len[u];
  }
  check_err_flag(b);
  SN_API_END
}

int64_t snmfnat_batch_out_len(const snmfnat_batch* b, int u) {
  if (!b || u < 0 || u >= b->n_utt) return -1;
  return b->out_len[u];
}
int64_t snmfnat_batch_total_hops(const snmfnat_batch* b) { return b ? b->NF : -1; }

int snmfnat_batch_get_stats(snmfnat_batch* b, snmfnat_batch_stats* out) {
  SN_API_BEGIN
  SN_REQUIRE(b && out, SNMFNAT_EINVAL, "NULL argument");
  SN_CUDA(cudaSetDevice(b->ctx->device));
  unsigned long long s[8];
  SN_CUDA(cudaMemcpyAsync(s, b->sb.stats.p, sizeof(s), cudaMemcpyDeviceToHost, b->ctx->stream));
  SN_CUDA(cudaStreamSynchronize(b->ctx->stream));
  std::memset(out, 0, sizeof(*out));
  out->hops = (int64_t)s[0]; out->h_iters = (int64_t)s[1]; out->w_iters = (int64_t)s[2]; out->gated_hops = (int64_t)s[3];
  out->w_solves = (int64_t)s[4]; out->w_atoms = (int64_t)s[5];
  // flops = sum_hops [ it_h*(4FR+10F) + it_w*(4*F*R_up*m_a + 12*F*m_a) + 0.7e6 ]; the W-solve term uses
  // the mean R_up of the run (sum over solves of it_w*R_up is approximated by w_iters * mean R_up).
  const Config& c = b->cfg;
  const double F = c.d.F, R = c.d.R, ma = c.d.m_a;
  const double mean_rup = s[4] ? (double)s[5] / (double)s[4] : 0.0;
  out->flops = (double)s[1] * (4.0 * F * R + 10.0 * F) + (double)s[2] * (4.0 * F * mean_rup * ma + 12.0 * F * ma) +
               (double)s[0] * 0.7e6;
  out->launches = b->launches_last;
  SN_API_END
}

int snmfnat_batch_set_mel(snmfnat_batch* b, const double* B_Mel_x, const double* B_Mel_d, int n1, const double* melmat) {
  SN_API_BEGIN
  SN_REQUIRE(b && B_Mel_x && B_Mel_d && n1 > 0, SNMFNAT_EINVAL, "bad argument");
  SN_REQUIRE(b->cfg.sc.mel_mode, SNMFNAT_EINVAL, "the batch was not created with B_sep_mode='Mel'");
  SN_REQUIRE(n1 <= 256, SNMFNAT_EUNSUPPORTED, "Mel mode supports up to 256 bands (got %d)", n1);
  SN_CUDA(cudaSetDevice(b->ctx->device));
  const Config& c = b->cfg;
  std::vector<double> M;
  if (!melmat) {  // init_buff.m:60-62: g.melmat = mel_matrix(p.fs, p.F_order, p.fftlength, 1, p.fs/2)'
    SN_REQUIRE(c.p.F_order == n1, SNMFNAT_EINVAL, "B_Mel has %d rows but p.F_order = %d", n1, c.p.F_order);
    M.resize((size_t)c.d.F * n1);
    mel_matrix_host(c.p.fs, n1, c.p.fftlength, 1.0, c.p.fs / 2.0, M.data());
    melmat = M.data();
  }
  b->sb.set_mel(b->ctx, n1, B_Mel_x, B_Mel_d, melmat);
  b->Ysep.alloc((size_t)std::max<long long>(b->NF, 1) * b->sb.LD1);
  SN_API_END
}

int snmfnat_batch_set_groups(snmfnat_batch* b, int n_groups) {
  SN_API_BEGIN
  SN_REQUIRE(b && n_groups >= 1 && n_groups <= 16, SNMFNAT_EINVAL, "n_groups must be in 1..16");
  b->n_groups = n_groups;
  SN_API_END
}

int snmfnat_batch_set_profile(snmfnat_batch* b, int on) {
  SN_API_BEGIN
  SN_REQUIRE(b, SNMFNAT_EINVAL, "NULL argument");
  b->profile = on != 0;
  SN_API_END
}

int snmfnat_batch_get_profile(snmfnat_batch* b, double* ms6, int64_t* counts6) {
  SN_API_BEGIN
  SN_REQUIRE(b && ms6, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(b->profile && b->ran, SNMFNAT_EINVAL, "profiling was not enabled for the last run");
  SN_CUDA(cudaSetDevice(b->ctx->device));
  SN_CUDA(cudaStreamSynchronize(b->ctx->stream));
  const size_t n = 3 * (size_t)b->max_hops + 3;
  double ms[6] = {0, 0, 0, 0, 0, 0};
  int64_t cnt[6] = {0, 0, 0, 0, 0, 0};
  auto el = [&](size_t a, size_t q) {
    float t = 0.f;
    SN_CUDA(cudaEventElapsedTime(&t, b->ev[a], b->ev[q]));
    return (double)t;
  };
  ms[0] = el(0, 1);
  cnt[0] = 3;
  for (int g = 0; g < b->max_hops; ++g) {
    const size_t e0 = 1 + 3 * (size_t)g;
    ms[1] += el(e0, e0 + 1);
    ms[2] += el(e0 + 1, e0 + 2);
    ms[3] += el(e0 + 2, e0 + 3);
    cnt[1]++; cnt[2]++; cnt[3]++;
  }
  ms[4] = el(n - 2, n - 1);
  cnt[4] = 3;
  ms[5] = el(0, n - 1);
  cnt[5] = b->launches_last;
  for (int i = 0; i < 6; ++i) {
    ms6[i] = ms[i];
    if (counts6) counts6[i] = cnt[i];
  }
  SN_API_END
}

int snmfnat_batch_get_trace(snmfnat_batch* b, int u, const char* what, double* buf, int64_t n) {
  SN_API_BEGIN
  SN_REQUIRE(b && what && buf, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(u >= 0 && u < b->n_utt, SNMFNAT_EINVAL, "utterance index out of range");
  SN_CUDA(cudaSetDevice(b->ctx->device));
  SN_CUDA(cudaStreamSynchronize(b->ctx->stream));
  const Config& c = b->cfg;
  const long long f0 = b->frame_base_u[u];
  const int nh = b->n_hops_u[u];
  const std::string w(what);
  auto rows = [&](const double* src, int ld, int width) {
    SN_REQUIRE(n >= (int64_t)nh * width, SNMFNAT_EINVAL, "buffer too small: need %lld doubles", (long long)nh * width);
    SN_CUDA(cudaMemcpy2D(buf, (size_t)width * sizeof(double), src + (size_t)f0 * ld, (size_t)ld * sizeof(double),
                         (size_t)width * sizeof(double), nh, cudaMemcpyDeviceToHost));
  };
  if (w == "Xm_tilde") { rows(b->Xt.p, c.d.LDF, c.d.F); }
  else if (w == "Ym") { rows(b->Ym.p, c.d.LDF, c.d.F); }
  else {
    SN_REQUIRE(b->trace && b->trA.p, SNMFNAT_EINVAL, "tracing was not enabled before the run");
    if (w == "A") rows(b->trA.p, c.d.R, c.d.R);
    else if (w == "Q") rows(b->trQ.p, c.d.LDF, c.d.F);
    else if (w == "G") rows(b->trG.p, c.d.LDF, c.d.F);
    else {
      int col = -1;
      if (w == "h_iters") col = 0; else if (w == "gated") col = 1; else if (w == "R_a_up") col = 2; else if (w == "w_iters") col = 3;
      SN_REQUIRE(col >= 0, SNMFNAT_EINVAL, "unknown trace field '%s'", what);
      SN_REQUIRE(n >= nh, SNMFNAT_EINVAL, "buffer too small");
      std::vector<int> tmp((size_t)nh * 4);
      SN_CUDA(cudaMemcpy(tmp.data(), b->trInfo.p + (size_t)f0 * 4, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost));
      for (int i = 0; i < nh; ++i) buf[i] = tmp[(size_t)i * 4 + col];
    }
  }
  SN_API_END
}

int snmfnat_batch_get_noise_basis(snmfnat_batch* b, int u, double* B_d) {
  SN_API_BEGIN
  SN_REQUIRE(b && B_d, SNMFNAT_EINVAL, "NULL argument");
  SN_REQUIRE(u >= 0 && u < b->n_utt, SNMFNAT_EINVAL, "utterance index out of range");
  SN_CUDA(cudaSetDevice(b->ctx->device));
  SN_CUDA(cudaStreamSynchronize(b->ctx->stream));
  const Config& c = b->cfg;
  const int s = b->slot_of[u];
  int sel = 0;
  SN_CUDA(cudaMemcpy(&sel, b->sb.bd_sel.p + s, sizeof(int), cudaMemcpyDeviceToHost));
  if (c.sc.mel_mode && b->sb.mel()) {  // the adapted basis is B_Mel_d (n1 x R_d)
    const double* srcm = (sel ? b->sb.BdM1.p : b->sb.BdM0.p) + (size_t)s * c.d.R_d * b->sb.LD1;
    download_basis(b->ctx, srcm, b->sb.n1, c.d.R_d, b->sb.LD1, B_d);
    return SNMFNAT_OK;
  }
  const double* src = (sel ? b->sb.Bd1.p : b->sb.Bd0.p) + (size_t)s * c.d.R_d * c.d.LDF;
  download_basis(b->ctx, src, c.d.F, c.d.R_d, c.d.LDF, B_d);
  SN_API_END
}

int snmfnat_enhance_batch(snmfnat_ctx* ctx, const snmfnat_params* p, const double* win_stft, const double* win_istft,
                          const double* B_x, const double* B_d, int n2, int n_utt, const int16_t* const* pcm,
                          const int64_t* len, const int32_t* chain_id, const double* h_init, const double* Ad_blk_init,
                          int64_t ad_stride, int16_t* const* out) {
  snmfnat_batch* b = nullptr;
  int rc = snmfnat_batch_create(ctx, p, win_stft, win_istft, B_x, B_d, n2, n_utt, len, chain_id, h_init, Ad_blk_init,
                                ad_stride, &b);
  if (rc) return rc;
  rc = snmfnat_batch_upload(b, pcm);
  if (!rc) rc = snmfnat_batch_run(b);
  if (!rc) rc = snmfnat_batch_download(b, out);
  snmfnat_batch_destroy(b);
  return rc;
}

}  // extern "C"
