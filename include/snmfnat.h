/*
 * snmfnat.h -- C ABI of libsnmfnat.so: the B200-native (sm_100a) implementation of the
 * SE_SNMF_NAT enhancement hot path.  This is the drop-in boundary: every entry point
 * replaces one MATLAB function of the reference (cited as file:line relative to the
 * reference root) and is what that function's MEX gateway (the .cpp files under mex/) binds.
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes; no C++/torch types.
 *   - all matrices are COLUMN-MAJOR IEEE doubles exactly as MATLAB stores them
 *     (mxGetPr of an F x R matrix), logical index vectors are uint8 (mxLogical).
 *   - every function returns 0 on success, a negative SNMFNAT_E* code otherwise; the
 *     message is available through snmfnat_last_error().  There is NO CPU fallback: if
 *     no sm_100 device is usable the call fails with SNMFNAT_ENODEVICE.
 *   - host buffers are owned by the caller; device memory by the library.
 *   - random numbers never originate inside the library: what the reference draws with
 *     rand() (sparse_nmf.m:112-134, init_buff.m:37-38) is passed in by the caller.
 *   - one host thread per context at a time.
 */
#ifndef SNMFNAT_H_
#define SNMFNAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNMFNAT_VERSION 100

enum {
  SNMFNAT_OK = 0,
  SNMFNAT_EINVAL = -1,       /* bad argument / shape */
  SNMFNAT_ENODEVICE = -2,    /* no usable sm_100 CUDA device */
  SNMFNAT_ECUDA = -3,        /* CUDA / cuFFT / NCCL runtime error */
  SNMFNAT_EUNSUPPORTED = -4, /* configuration the reference itself cannot run (Splice>0, blk_len_sep>1) or not implemented */
  SNMFNAT_ENUMERIC = -5,     /* reference would hit a dimension mismatch (all-zero activation row, bnmf_sep_event_RT_IS16.m:292) */
  SNMFNAT_ENOMEM = -6
};

#define SNMFNAT_MAX_CLASSES 8

/* cost function selector, sparse_nmf.m:95-110 */
enum { SNMFNAT_CF_IS = 0, SNMFNAT_CF_KL = 1, SNMFNAT_CF_ED = 2, SNMFNAT_CF_BETA = 3 };
enum { SNMFNAT_ENH_MMSE = 0, SNMFNAT_ENH_WIENER = 1 };
enum { SNMFNAT_SEP_DFT = 0, SNMFNAT_SEP_MEL = 1 };

/* Flat copy of the fields of `global p` that the hot path reads
 * (settings/initial_setting_SNMF_NAT.m:1-149; list in SURVEY.md 5.1). */
typedef struct snmfnat_params {
  int32_t fs, framelength, frameshift, fftlength, delay;
  int32_t blk_len_sep, blk_hop_sep, Splice;
  int32_t EVENT_NUM, NOISE_NUM;
  int32_t EVENT_RANK[SNMFNAT_MAX_CLASSES], NOISE_RANK[SNMFNAT_MAX_CLASSES]; /* 1-based like the reference */
  int32_t R_x, R_d, R_a, m_a, init_N_len, adapt_train_N;
  int32_t blk_sparse, P_len_k, P_len_l, blk_gap;
  int32_t DCbin, DCbin_back, F_order;
  int32_t B_sep_mode, MelConv;
  int32_t cf, max_iter, cost_check, basis_update_N, basis_update_E;
  int32_t ENHANCE_METHOD;
  int32_t reserved_i[7];
  double overlapscale, pow, nonzerofloor;
  double overlap_m_a, Ar_up, alpha_p, preemph;
  double beta_div;              /* used when cf == SNMFNAT_CF_BETA */
  double sparsity, conv_eps;
  double alpha_eta, alpha_d, beta, beta_max;
  double sparsity_mdi, conv_eps_mdi;
  double reserved_d[6];
} snmfnat_params;

/* Options of one sparse_nmf call (the optional fields of p, sparse_nmf.m:79-164,260). */
typedef struct snmfnat_nmf_opts {
  int32_t max_iter;     /* p.max_iter, default 100 */
  int32_t cf;           /* SNMFNAT_CF_* */
  int32_t cost_check;   /* p.cost_check (no default in the reference) */
  int32_t sparsity_rows, sparsity_cols; /* shape of `sparsity`: 1x1, r x 1 or r x n */
  int32_t precision;    /* 0 = float64 (online parity path); 1 = TF32 tensor-core training path (fp32 state) */
  double beta_div;      /* p.beta when cf == SNMFNAT_CF_BETA */
  double conv_eps;      /* p.conv_eps */
} snmfnat_nmf_opts;

typedef struct snmfnat_ctx snmfnat_ctx;       /* one CUDA device + stream + cuFFT plans */
typedef struct snmfnat_stream snmfnat_stream; /* device-resident struct g of one audio stream */
typedef struct snmfnat_batch snmfnat_batch;   /* a batch of utterances advanced hop-by-hop in lock step */
typedef struct snmfnat_train snmfnat_train;   /* frame-sharded dictionary training state */

/* ---- context ------------------------------------------------------------------------- */
int snmfnat_version(void);
int snmfnat_ctx_create(int device, snmfnat_ctx** out);
int snmfnat_ctx_destroy(snmfnat_ctx* ctx);
int snmfnat_ctx_sync(snmfnat_ctx* ctx);
/* The CUDA stream every kernel of this context is launched on (a cudaStream_t), so that a
 * caller can bracket calls with its own CUDA events. */
void* snmfnat_ctx_cuda_stream(snmfnat_ctx* ctx);
/* Message of the last failing call on this thread (ctx may be NULL). */
const char* snmfnat_last_error(const snmfnat_ctx* ctx);
/* Kernels launched by this context so far (our own kernels; cuFFT launches not counted). */
int64_t snmfnat_ctx_launch_count(const snmfnat_ctx* ctx);

/* Fill *p with the shipped configuration, settings/initial_setting_SNMF_NAT.m:1-149. */
void snmfnat_params_default(snmfnat_params* p);

/* ---- L1 numeric kernels ----------------------------------------------------------- */

/* [w,h,objective] = sparse_nmf(v,p)                     src/sparse_nmf.m:1,71-292
 *   v F x n; init_w F x r (required: RNG stays with the caller); init_h r x n (required);
 *   w_ind/h_ind r logicals (NULL = all true); sparsity per opts->sparsity_rows/cols.
 *   Outputs: w F x r, h r x n, div/cost arrays of length opts->max_iter (first *iters valid;
 *   all-zero when !cost_check), *iters = executed iterations. */
int snmfnat_sparse_nmf(snmfnat_ctx* ctx, const double* v, int F, int n, int r,
                       const snmfnat_nmf_opts* opts, const double* sparsity,
                       const double* init_w, const double* init_h,
                       const uint8_t* w_ind, const uint8_t* h_ind,
                       double* w, double* h, double* div, double* cost, int* iters);

/* [v_MDI,h,objective] = snmf_mdi(v,Dm,p)                src/snmf_mdi.m:1,175,251-255,297-303
 * and snmf_mdi_Sm (soft != 0)                              src/snmf_mdi_Sm.m:175,252-260,307-309
 *   mask F x n doubles (0/1 for Dm, [0,1] for Sm); opts->conv_eps / sparsity carry
 *   p.conv_eps_mdi / p.sparsity_mdi. */
int snmfnat_snmf_mdi(snmfnat_ctx* ctx, const double* v, const double* mask, int soft, int F, int n, int r,
                     const snmfnat_nmf_opts* opts, const double* sparsity,
                     const double* init_w, const double* init_h,
                     const uint8_t* w_ind, const uint8_t* h_ind,
                     double* v_mdi, double* h, double* div, double* cost, int* iters);

/* B_a = DNMF_adapt(Y,D,B,p)                                src/DNMF_adapt.m:1-21
 *   Y,D F x n; B F x (R_x+R_d); h_init (R_x+R_d) x n (the rand(r,n) of the inner H-solve);
 *   output B_a F x R_d. */
int snmfnat_dnmf_adapt(snmfnat_ctx* ctx, const double* Y, const double* D, const double* B,
                       int F, int n, int R_x, int R_d, const snmfnat_nmf_opts* opts, const double* sparsity,
                       const double* h_init, double* B_a);

/* [S_mag,S_phase] = stft_fft(s,sz,shift,fftlen,DCbin,win,preemph)   src/stft_fft.m:1-37
 *   outputs (fftlen/2+1) x floor(len/shift), trailing columns zero like the reference. */
int snmfnat_stft_fft(snmfnat_ctx* ctx, const double* s, int64_t len, int sz, int shift, int fftlen, int DCbin,
                     const double* win, double preemph, double* S_mag, double* S_phase);

/* s_buff = synth_ifft_buff(TF_mag,TF_phase,sz,fftlen,win,preemph,DCbin_back,pow)   src/synth_ifft_buff.m:1-33
 *   TF_mag/TF_phase freq_num x frame_num with freq_num == fftlen/2+1 or fftlen; s_buff sz x frame_num. */
int snmfnat_synth_ifft_buff(snmfnat_ctx* ctx, const double* TF_mag, const double* TF_phase, int freq_num,
                            int frame_num, int sz, int fftlen, const double* win, double preemph,
                            int DCbin_back, double pow_, double* s_buff);

/* [Q,r_blk_out] = blk_sparse(X,D,r_blk,l,p)               src/blk_sparse.m:1-37
 *   X,D K x 1; r_blk K x P_len_l; outputs Q K x 1, r_blk_out K x P_len_l. */
int snmfnat_blk_sparse(snmfnat_ctx* ctx, const double* X, const double* D, const double* r_blk, int K, int l,
                       const snmfnat_params* p, double* Q, double* r_blk_out);

/* ---- L2 per-hop entry (latency path) --------------------------------------------------- */

/* g = init_buff(B_Mel_x,B_Mel_d,B_DFT_x,B_DFT_d,p)        src/init_buff.m:1-62
 *   n1 = rows of the "Mel" slot (== n2 in DFT mode, filewise_run_IS16.m:46-51), n2 = fftlength/2+1.
 *   win_stft/win_istft: framelength doubles (p.win_STFT / p.win_ISTFT).
 *   Ad_blk_init R_a x m_a replaces rand(p.R_a,p.m_a) (:38); A_d_init R_d (may be NULL) replaces :37. */
int snmfnat_stream_create(snmfnat_ctx* ctx, const snmfnat_params* p, const double* win_stft,
                          const double* win_istft, const double* B_Mel_x, const double* B_Mel_d, int n1,
                          const double* B_DFT_x, const double* B_DFT_d, int n2, const double* Ad_blk_init,
                          const double* A_d_init, snmfnat_stream** out);
int snmfnat_stream_destroy(snmfnat_stream* s);

/* [x_hat_i,d_hat_i,x_tilde,g] = bnmf_sep_event_RT_IS16(y,l,g,p)   src/bnmf_sep_event_RT_IS16.m:1-423
 *   y framelength doubles; l 1-based hop index; h_init (R_x+R_d) = rand(r,1) after rand('seed',..).
 *   x_tilde framelength doubles; x_hat_i (EVENT_NUM x framelength) / d_hat_i (NOISE_NUM x framelength)
 *   may be NULL (callers discard them: filewise_run_IS16.m:142).  Synchronous. */
int snmfnat_stream_step(snmfnat_stream* s, const double* y, int l, const double* h_init, double* x_tilde,
                        double* x_hat_i, double* d_hat_i);

/* Read / write one field of g by its reference name ("B_DFT_d", "Ad_blk", "lambda_d_blk", "r_blk",
 * "lambda_dav", "Xm_tilde", "Ym", "Yp", "B_Mel_d", "B_DFT_x", "B_Mel_x", "update_switch", "A" (last
 * activations), "Q", "G", "Xm_hat", "Dm_hat", "stats" = {h_iters, gated, R_a_up, w_iters, h_cost}).
 * n = number of doubles of the buffer; layouts are MATLAB's (history matrices oldest column first). */
int snmfnat_stream_get(snmfnat_stream* s, const char* field, double* buf, int64_t n);
int snmfnat_stream_set(snmfnat_stream* s, const char* field, const double* buf, int64_t n);

/* ---- L3 whole-batch entry (throughput path) ---------------------------------------------- */

/* The hop loops of filewise_run_IS16.m:86-169 (chain_id == NULL: every utterance independent) or of
 * src/NTF_sep_event_RT.m:28-139 (chain_id[u] >= 0: utterances with equal id form a chain in index order,
 * the adapted noise basis is carried file to file like B_D_u.mat), run on the device for n_utt
 * utterances in lock step.  len[u] = int16 samples of utterance u (header already stripped).
 *   B_x n2 x R_x, B_d n2 x R_d (DFT mode; Mel slots take the same matrices, filewise_run_IS16.m:46-51).
 *   h_init (R_x+R_d) shared by all hops (sparse_nmf.m:112-114); Ad_blk_init R_a x m_a per utterance
 *   (ad_stride doubles apart; 0 = one matrix shared by all). */
int snmfnat_batch_create(snmfnat_ctx* ctx, const snmfnat_params* p, const double* win_stft,
                         const double* win_istft, const double* B_x, const double* B_d, int n2, int n_utt,
                         const int64_t* len, const int32_t* chain_id, const double* h_init,
                         const double* Ad_blk_init, int64_t ad_stride, snmfnat_batch** out);
int snmfnat_batch_destroy(snmfnat_batch* b);
/* Host -> device copy of the PCM of every utterance (pcm[u] has len[u] samples). */
int snmfnat_batch_upload(snmfnat_batch* b, const int16_t* const* pcm);
/* Same, from one contiguous (preferably pinned) buffer holding the utterances back to back.  ASYNCHRONOUS on the context
 * stream, unlike snmfnat_batch_upload: the copies are only queued when the call returns, so pcm_packed must stay valid
 * and unmodified until the next synchronising call on this context (snmfnat_ctx_sync, snmfnat_batch_download*, ...).
 * A caller that refills or frees the buffer earlier feeds undefined PCM into the run. */
int snmfnat_batch_upload_packed(snmfnat_batch* b, const int16_t* pcm_packed);
/* Reset the per-stream state to init_buff and enhance every utterance; inputs and outputs stay in HBM.
 * Asynchronous on the context stream; snmfnat_ctx_sync() or a download waits for it. */
int snmfnat_batch_run(snmfnat_batch* b);
/* Device -> host copy of the enhanced int16 signals; out[u] must hold snmfnat_batch_out_len(b,u) samples
 * ((floor(len/frameshift)+1)*frameshift, filewise_run_IS16.m:146,162-165). */
/* B_sep_mode = 'Mel' (bnmf_sep_event_RT_IS16.m:107-119,165-211,295-319): the separation and the noise-basis
 * adaptation run on n1 Mel bands.  B_Mel_x n1 x R_x, B_Mel_d n1 x R_d (the B_Mel_sub of the shipped basis files); melmat is the
 * n2 x n1 matrix src/mel_matrix.m returns, or NULL to have the library build it the way init_buff.m:60-62 does.
 * Must be called once, before snmfnat_batch_run, when p.B_sep_mode == SNMFNAT_SEP_MEL (MelConv = 1 only);
 * snmfnat_batch_get_noise_basis then returns the adapted B_Mel_d (n1 x R_d). */
int snmfnat_batch_set_mel(snmfnat_batch* b, const double* B_Mel_x, const double* B_Mel_d, int n1, const double* melmat);
/* M = mel_matrix(fs, NbCh, Nfft, warp, fhigh), src/mel_matrix.m:9-38: (Nfft/2+1) x NbCh, column-major.
 * warp <= 0 selects 1, fhigh <= 0 selects fs/2 (the defaults of the reference). */
int snmfnat_mel_matrix(int fs, int NbCh, int Nfft, double warp, double fhigh, double* M);
/* Training features of run_basis_train.m:63,70-78 / run_basis_DNMF.m:15,24,33 / run_basis_DNMF_Mel.m:16-27:
 * out = S_mag.^pow + floor (F x T), or melmat' * that (n1 x T) when melmat (F x n1, column-major) is given. */
int snmfnat_tf_features(snmfnat_ctx* ctx, const double* S_mag, int F, int64_t T, double pow_, double floor_,
                        const double* melmat, int n1, double* out);
/* [C, A] = GIST_NTF(p, B, S_mag)  (src/GIST_NTF.m:1-160; cost_check = -1) and GIST_NTF_C (src/GIST_NTF_C.m; cost_check =
 * p.cost_check): KL tensor factorisation S(h,n,m) ~ sum_k C(h,k) B(n,k) A(m,k) with only the channel gains updated.
 * S_mag Channel x N x M and B N x K column-major; C_init = rand(Channel, K) drawn by the host (:14); A M x K or NULL for
 * the reference's ones(M, K) (:16).  C_out Channel x K; div / cost hold max_iter doubles (may be NULL). */
int snmfnat_gist_ntf(snmfnat_ctx* ctx, const double* S_mag, int Channel, int N, int M, const double* B, int K,
                     const double* C_init, const double* A, double sparsity, double flr, int max_iter, double conv_eps,
                     int cost_check, double* C_out, double* div, double* cost, int* iters);
/* Scheduling knob, no effect on results: the slots are split into n_groups interleaved groups whose per-hop kernels
 * run on separate CUDA streams, so that the tail of one group's kernel overlaps the next kernel of another group.
 * Default 3 (or the SNMFNAT_GROUPS environment variable). */
int snmfnat_batch_set_groups(snmfnat_batch* b, int n_groups);
int snmfnat_batch_download(snmfnat_batch* b, int16_t* const* out);
int snmfnat_batch_download_packed(snmfnat_batch* b, int16_t* out_packed);
int64_t snmfnat_batch_out_len(const snmfnat_batch* b, int u);
int64_t snmfnat_batch_total_hops(const snmfnat_batch* b);

typedef struct snmfnat_batch_stats {
  int64_t hops;            /* hops executed (all utterances, flush hops included) */
  int64_t h_iters;         /* MU iterations of all H-solves */
  int64_t w_iters;         /* MU iterations of all W-solves */
  int64_t gated_hops;      /* hops whose adaptation gate fired */
  int64_t w_solves;        /* W-solves executed (R_a_up > 0) */
  int64_t w_atoms;         /* sum of R_a_up over W-solves */
  double flops;            /* algorithmic flops, SURVEY.md 8(d) formula with the actual counts */
  int64_t launches;        /* kernels launched by the last snmfnat_batch_run */
  int64_t reserved[4];
} snmfnat_batch_stats;
int snmfnat_batch_get_stats(snmfnat_batch* b, snmfnat_batch_stats* out);
/* Per-hop trace of utterance u for parity tests: what[] in {"A","Xm_tilde","Q","G","h_iters","gated",
 * "R_a_up","w_iters"}; only available when tracing was enabled before the run. */
int snmfnat_batch_enable_trace(snmfnat_batch* b, int on);
int snmfnat_batch_get_trace(snmfnat_batch* b, int u, const char* what, double* buf, int64_t n);
/* Per-kernel-class device times of the last run, measured with CUDA events recorded on the context stream
 * around every launch (enable before the run).  ms[0..5] = STFT (framing + cuFFT + epilogue), H-solve,
 * gain/blk_sparse, W-solve, ISTFT + overlap-add, whole run; counts[0..5] = launches of each class. */
int snmfnat_batch_set_profile(snmfnat_batch* b, int on);
int snmfnat_batch_get_profile(snmfnat_batch* b, double* ms6, int64_t* counts6);
/* Final adapted noise basis of utterance u (n2 x R_d), i.e. g.B_DFT_d after the last hop. */
int snmfnat_batch_get_noise_basis(snmfnat_batch* b, int u, double* B_d);

/* One-call convenience: create + upload + run + download + destroy. */
int snmfnat_enhance_batch(snmfnat_ctx* ctx, const snmfnat_params* p, const double* win_stft,
                          const double* win_istft, const double* B_x, const double* B_d, int n2, int n_utt,
                          const int16_t* const* pcm, const int64_t* len, const int32_t* chain_id,
                          const double* h_init, const double* Ad_blk_init, int64_t ad_stride,
                          int16_t* const* out);

/* The same on n_dev GPUs of one node from ONE host process (what a single MATLAB / Octave process binds
 * for Do_MultiBatch_IS16_20160324_CHiME4.m:202-208 / run_ntf_sep_RT.m:9-41): utterances, or chains of utterances, are
 * split over devices[0..n_dev) longest-processing-time first; one host thread and one context per device; no collective.
 * Arguments as snmfnat_enhance_batch (indices of pcm / len / chain_id / Ad_blk_init / out are corpus indices). */
int snmfnat_enhance_batch_multi(const int* devices, int n_dev, const snmfnat_params* p, const double* win_stft,
                                const double* win_istft, const double* B_x, const double* B_d, int n2, int n_utt,
                                const int16_t* const* pcm, const int64_t* len, const int32_t* chain_id,
                                const double* h_init, const double* Ad_blk_init, int64_t ad_stride,
                                int16_t* const* out);

/* ---- offline dictionary training (run_basis_train.m:80-91,112-116) ---------------------- */

/* Frame-sharded sparse_nmf with W and H both updated (sparse_nmf.m:186-286), KL, on T_local frames held
 * by this rank.  V F x T_local (float32, column-major, device or host per `v_on_device`), init_w F x K,
 * init_h K x T_local.  The W-update accumulators are all-reduced across ranks when a communicator is
 * attached (snmfnat_train_attach_nccl). */
int snmfnat_train_create(snmfnat_ctx* ctx, int F, int K, int64_t T_local, double sparsity, int precision,
                         snmfnat_train** out);
int snmfnat_train_destroy(snmfnat_train* t);
/* nccl_unique_id: the 128 bytes of an ncclUniqueId created on rank 0 and broadcast by the host side. */
int snmfnat_train_nccl_unique_id(void* id128);
int snmfnat_train_attach_nccl(snmfnat_train* t, const void* nccl_unique_id, int rank, int world);
int snmfnat_train_set_data(snmfnat_train* t, const float* V, int v_on_device, const float* init_w,
                           const float* init_h, int h_on_device);
/* Device pointers of the resident arrays (float32) so that a host side can fill them in place:
 * "V" [T_local][ldv] (frame-major, ldv >= F), "H" [T_local][Kp] (frame-major, Kp >= K, padding must stay 0),
 * "W_init" [K][F] (the init_w staging read by snmfnat_train_reset), "W" [K][F] (current dictionary). */
void* snmfnat_train_dev_ptr(snmfnat_train* t, const char* which);
int snmfnat_train_get_layout(snmfnat_train* t, int* ldv, int* kp);
/* Call after filling "V" in place: rebuilds the library's bin-major copy of V. */
int snmfnat_train_commit_v(snmfnat_train* t);
/* Re-run the init of sparse_nmf.m:157-160 (normalise W_init into W, rescale H) on the resident data. */
int snmfnat_train_reset(snmfnat_train* t);
/* n_iters MU iterations (H-update, W-update + all-reduce), sparse_nmf.m:186-244.  div / cost (n_iters doubles each,
 * sparse_nmf.m:250,261) may be NULL; asking for them costs one stream synchronisation per iteration. */
int snmfnat_train_iterate(snmfnat_train* t, int n_iters, double* div, double* cost);
/* The whole loop of sparse_nmf.m:186-286 with its stop rule (:273-283): at most max_iter iterations, stops after
 * iteration it > 1 when |cost - last_cost| / last_cost < conv_eps.  div / cost hold max_iter doubles (may be NULL),
 * *iters = executed iterations. */
int snmfnat_train_run(snmfnat_train* t, int max_iter, double conv_eps, double* div, double* cost, int* iters);
/* The all-reduced W-update accumulators of the last iteration: g = (V ./ (W*H)) * H' (F x K, sparse_nmf.m:217) and
 * hs = sum(H,2) (K); either may be NULL. */
int snmfnat_train_get_acc(snmfnat_train* t, float* g, float* hs);
int snmfnat_train_get_w(snmfnat_train* t, float* w);
int snmfnat_train_get_h(snmfnat_train* t, float* h, int64_t t0, int64_t count);

#ifdef __cplusplus
}
#endif
#endif /* SNMFNAT_H_ */
