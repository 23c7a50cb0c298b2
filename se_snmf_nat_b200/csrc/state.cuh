// Owner of the device-resident per-slot state (the struct g of src/init_buff.m for S streams) and the
// translation of snmfnat_params into kernel-side dimension / scalar blocks.
#pragma once
#include "online.cuh"
#include "stft.cuh"

namespace snmfnat {

struct Config {
  snmfnat_params p;
  OnlineDims d;
  OnlineScalars sc;
  StftGeom g;
};
// Validates p against what the IS16 frame function supports and fills the kernel-side blocks.
void make_config(snmfnat_ctx* ctx, const snmfnat_params& p, int n2, Config& cfg);

struct SlotBuffers {
  int S = 0;
  OnlineDims d{};
  DevBuf<double> Bx, Bd_fix, Bd0, Bd1, Ad_blk, Ad_init, lam_blk, r_blk, lambda_dav, Xm_tilde_prev;
  DevBuf<double> A, Xhat, Dhat, Q, G, h_cost, h_init, win_stft, win_istft;
  DevBuf<int> rblk_pos, bd_sel, ring_head, update_switch, h_iters, gated, do_update, n_up, idx_up, idx_rem, w_iters, err_flag;
  DevBuf<int> l_offset, n_hops, first_utt_dev;
  int n_utt_ad = 0;   // utterances held in Ad_init
  DevBuf<long long> frame_base;
  DevBuf<unsigned long long> stats;
  DevBuf<double> ms_colstat;   // norms / sums of the stream-invariant columns (multi-stream H-solve), when supported
  DevBuf<int> ms_perm, ms_perm_step, ms_ticket;   // launch order of the next hop per slot group (SlotState::ms_perm)
  DevBuf<int> ws_perm, ws_perm_step, w_last;      // launch order of the W-solve (SlotState::ws_perm)
  DevBuf<double> semi_w;       // private copies of the atoms a semi-supervised separation solve updates (SlotState::semi_w)

  // Mel separation mode: Mel-sized bases / history / reconstructions next to the DFT-domain state
  int n1 = 0, LD1 = 0;
  DevBuf<double> melM, BxM, BdM_fix, BdM0, BdM1, lam_blk_mel, XhatM, DhatM;
  bool mel() const { return n1 > 0; }
  // B_Mel_x n1 x R_x, B_Mel_d n1 x R_d, melmat n2 x n1 (host column-major, i.e. the matrix mel_matrix.m returns)
  void set_mel(snmfnat_ctx* ctx, int n1, const double* B_Mel_x, const double* B_Mel_d, const double* melmat);
  OnlineDims dims_mel() const;
  SlotState view_mel() const;

  void alloc(int S, const OnlineDims& d);
  // bases are host column-major F x R doubles
  void set_bases(snmfnat_ctx* ctx, const double* B_x, const double* B_d);
  // Ad_blk_init: R_a x m_a column-major per UTTERANCE, `stride` doubles apart (0 = shared by all); first_utt[s] = the
  // utterance slot s starts with (later members of a chain are installed by chain_boundary)
  void set_ad_init(snmfnat_ctx* ctx, const double* Ad_blk_init, int64_t stride, int n_utt, const std::vector<int>& first_utt);
  // A slot moves on to the next file of its chain (src/NTF_sep_event_RT.m:28-46: init_buff with the bases loaded from
  // B_D_u.mat): everything of init_buff is reset EXCEPT the adapted noise basis; events = {slot, utt, step0, n_hops}
  void chain_boundary(snmfnat_ctx* ctx, const int* events_dev, int n_events);
  // reset every slot to init_buff (src/init_buff.m:17-62): zero histories, Bd <- B_d, Ad_blk <- init
  void reset(snmfnat_ctx* ctx);
  SlotState view() const;
};

// host column-major F x R  <->  device [R][LDF]
void upload_basis(snmfnat_ctx* ctx, const double* host, int F, int R, int LDF, double* dev);
void download_basis(snmfnat_ctx* ctx, const double* dev, int F, int R, int LDF, double* host);

}  // namespace snmfnat
