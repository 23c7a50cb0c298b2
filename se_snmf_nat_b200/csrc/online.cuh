// Device-resident state of a set of audio streams ("slots") and the launchers of the per-hop kernels.
// One slot = the struct g of src/init_buff.m:17-62 for one stream, laid out for the kernels:
//   - every F-long vector / basis column has padded leading dimension LDF (multiple of 8 doubles)
//   - history matrices (lambda_d_blk, Ad_blk, r_blk) are ring buffers indexed by time slot
//   - the adapted noise basis is double-buffered (the reference re-orders its columns on every update,
//     bnmf_sep_event_RT_IS16.m:336)
#pragma once
#include "common.cuh"

namespace snmfnat {

struct OnlineDims {
  int F;      // n2 = fftlength/2+1 rows of the DFT-domain vectors
  int LDF;    // padded F
  int R_x, R_d, R, R_a, m_a, P_len_l;
  // the slots a launch covers: slot0 + i * slot_stride, i = 0 .. n_active-1 (interleaved groups run on separate CUDA
  // streams so that the tail of one group's kernel overlaps the next kernel of another group)
  int slot0 = 0, slot_stride = 1;
  // atoms [upd0, upd1) that the separation solve itself updates (p.basis_update_N / _E, bnmf_sep_event_RT_IS16.m:125-139);
  // empty for the supervised solve of the shipped settings
  int upd0 = 0, upd1 = 0;
};

// Scalars of p used inside the kernels.
struct OnlineScalars {
  double flr;            // p.nonzerofloor (also sparse_nmf's flr = 1e-9, sparse_nmf.m:166)
  double sparsity, conv_eps;
  int max_iter, cost_check;
  int DCbin, init_N_len, adapt_train_N, blk_sparse, P_len_k, P_len_l, blk_gap;
  double alpha_p, alpha_eta, alpha_d, beta, beta_max, Ar_up;
  int enhance_method;
  int update_period;     // floor(p.overlap_m_a * p.m_a), bnmf_sep_event_RT_IS16.m:293
  int mel_mode;          // B_sep_mode = 'Mel': lambda_dav of the first hop comes from the Mel -> DFT image (:205-211,223-225)
};

// Pointers into the per-slot state arrays (slot-major).
struct SlotState {
  int S;                        // number of slots
  const double* Bx;             // [R_x][LDF] speech basis, shared by all slots (never changes in DFT mode)
  const double* Bd_fix;         // [R_d][LDF] the "B_Mel_d" slot in DFT mode: initial noise basis per slot? shared
  size_t bdfix_stride;          // 0 when shared by all slots, R_d*LDF when per slot (chains)
  double* Bd[2];                // [S][R_d][LDF] ping-pong adapted noise basis
  int* bd_sel;                  // [S] current buffer
  double* Ad_blk;               // [S][m_a][R_a]   ring (time-slot major)
  double* lam_blk;              // [S][m_a][LDF]   ring
  int* ring_head;               // [S] oldest time slot == next to overwrite
  double* r_blk;                // [S][P_len_l][LDF] ring
  int* rblk_pos;                // [S] slot that receives the next column (the oldest one)
  double* lambda_dav;           // [S][LDF]
  double* Xm_tilde_prev;        // [S][LDF]
  int* update_switch;           // [S]
  // per-hop scratch
  double* A;                    // [S][R]   activations of the H-solve (the reference's A)
  double* Xhat;                 // [S][LDF] B_x*A_x   (sum over event classes)
  double* Dhat;                 // [S][LDF] B_d*A_d
  double* Q;                    // [S][LDF]
  double* G;                    // [S][LDF]
  int* h_iters;                 // [S]
  double* h_cost;               // [S]
  int* gated;                   // [S]
  int* do_update;               // [S] W-solve requested this hop
  int* n_up;                    // [S]
  int* idx_up;                  // [S][R_a]
  int* idx_rem;                 // [S][R_a]
  int* w_iters;                 // [S]
  int* err_flag;                // [1] set when the reference would hit a dimension mismatch
  // hop bookkeeping
  const int* l_offset;          // [S] local hop index l = g_step + 1 - l_offset[s]  (1-based)
  const int* n_hops;            // [S] hops of the current utterance of this slot (slot inactive when l > n_hops)
  const long long* frame_base;  // [S] frame index of (slot, g_step = 0)
  // accumulators
  unsigned long long* stats;    // [8]: hops, h_iters, w_iters, gated, w_solves, w_atoms
  // multi-stream H-solve (online_ms.cu): norms / sums of the stream-invariant columns, optional launch order of the slots
  const double* ms_colstat = nullptr;
  // launch order of the next hop, written by the last CTA of gain_kernel: positions of the group's slots sorted by the
  // iteration count of this hop, longest first ([16][ms_perm_stride]); ms_perm_step[group] = the step it is valid for
  int* ms_perm = nullptr;
  int* ms_perm_step = nullptr;
  int* ms_ticket = nullptr;
  int ms_perm_stride = 0;
  // launch order of THIS hop's W-solve, written by the same CTA: gated slots first, longest expected solve first (passes of
  // the slot's previous solve x atom tiles of this one), so that the launch does not end in a tail of a few long solves
  int* ws_perm = nullptr;        // [16][ms_perm_stride]
  int* ws_perm_step = nullptr;   // [16] the step it is valid for
  int* w_last = nullptr;         // [S] passes of the slot's last W-solve
  // semi-supervised separation solve (online_semi.cu): normalised private copy of the atoms the solve updates
  double* semi_w = nullptr;      // [S][upd1-upd0][LDF]
};

// Per-frame arrays shared by the STFT, the solvers and the ISTFT.
struct FrameArrays {
  const double* Ym;   // [NF][LDF]  |Y|^pow (+floor, DC zeroed)
  double* Xt;         // [NF][LDF]  enhanced spectrum G.*Ym
};

// Optional per-hop trace for parity tests ([NF] rows).
struct TraceArrays {
  double* A;       // [NF][R]
  double* Q;       // [NF][LDF]
  double* G;       // [NF][LDF]
  int* info;       // [NF][4]  h_iters, gated, n_up, w_iters
};

// ---- launchers (online_kernels.cu) ----
// H-solve + reconstruction for slots [0, n_active) at global step g_step.
void launch_hsolve(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                   const FrameArrays& fr, const double* h_init, int n_active, int g_step);
// blk_sparse + gain + adaptation gate/history for slots [0, n_active).
// n_active_next >= 0: also sort the slots of the launch for the next hop's multi-stream H-solve (see SlotState::ms_perm)
void launch_gain(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                 const FrameArrays& fr, const TraceArrays* tr, int n_active, int g_step, int n_active_next = -1);
// W-solve (noise-basis adaptation) for the slots whose do_update flag is set.
void launch_wsolve(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                   const TraceArrays* tr, int n_active, int g_step);
// fast paths for the shipped geometry (online_fast.cu); the launchers above dispatch to them when supported
bool hsolve_fast_supported(snmfnat_ctx* ctx, const OnlineDims& d);
void launch_hsolve_fast(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                        const FrameArrays& fr, const double* h_init, int n_active, int g_step);
// multi-stream H-solve (online_ms.cu): S streams per 8-CTA cluster, stream-invariant columns as FP64 tensor-core fragments
bool hsolve_ms_supported(snmfnat_ctx* ctx, const OnlineDims& d);
int hsolve_ms_streams();
void launch_ms_colstat(snmfnat_ctx* ctx, const OnlineDims& d, const double* Bx, const double* Bd_fix, double* colstat);
void launch_hsolve_ms(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                      const FrameArrays& fr, const double* h_init, int n_active, int g_step);
// semi-supervised separation solve (online_semi.cu): d.upd0 < d.upd1, one CTA per stream
void launch_hsolve_semi(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                        const FrameArrays& fr, const double* h_init, int n_active, int g_step);
// SNMFNAT_HSOLVE=ms forces the multi-stream kernel for any number of active streams, =single disables it; by default it
// runs when at least hsolve_ms_streams() streams are active
int hsolve_ms_mode();
bool wsolve_fast_supported(snmfnat_ctx* ctx, const OnlineDims& d);
void launch_wsolve_fast(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                        const TraceArrays* tr, int n_active, int g_step);
// Mel separation mode (mel.cu); M is the filterbank [n1][LDF] (band-major)
void mel_matrix_host(int fs, int NbCh, int Nfft, double warp, double fhigh, double* M);
void launch_mel_project(snmfnat_ctx* ctx, const double* M, int n1, int LD1, int F, int LDF, const double* Ym, long long NF,
                        double* Ysep);
void launch_mel_post(snmfnat_ctx* ctx, const OnlineDims& d, const SlotState& st, const double* M, int n1, int LD1,
                     const double* XhatM, const double* DhatM, const double* Ysep, int n_active, int g_step);
void launch_mel_hist(snmfnat_ctx* ctx, const OnlineDims& d, const SlotState& st, const double* M, int n1, int LD1,
                     double* lam_blk_mel, int n_active, int g_step);
// SNMFNAT_FORCE_GENERIC=1 in the environment disables the fast paths (used by the parity tests)
bool force_generic();
// shared-memory footprints (bytes) so that callers can reject configurations that do not fit
size_t hsolve_smem_bytes(const OnlineDims& d);
size_t wsolve_smem_bytes(const OnlineDims& d);

}  // namespace snmfnat
