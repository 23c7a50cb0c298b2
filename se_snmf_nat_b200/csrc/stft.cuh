// STFT / ISTFT kernels around cuFFT (float64): framing + window + pre-emphasis, power/phase epilogue,
// spectrum synthesis, window + overlap-add + int16 conversion.
#pragma once
#include "common.cuh"

namespace snmfnat {

struct StftGeom {
  int sz, shift, fftlen, half;  // framelength, frameshift, fftlength, fftlength/2+1
  int LDF;                      // padded half
  int delay;
  double preemph, pow_, flr, overlapscale;
  int DCbin, DCbin_back;
};

// Utterance tables on the device (one entry per utterance).
struct UttTables {
  const long long* pcm_off;     // offset of the utterance in the packed PCM buffer
  const long long* len;         // samples
  const long long* frame_base;  // index of its first frame (hop l = 1) in the frame arrays
  const int* n_hops;            // hops = floor(len/shift) + delay + 1
  const long long* out_off;     // offset in the packed output buffer
  int n_utt, max_hops;
};

// int16 PCM -> windowed zero-padded frames [NF][fftlen]   (filewise_run_IS16.m:102-123 queue + IS16 fn :66-69)
void launch_frame_pcm(snmfnat_ctx* ctx, const StftGeom& g, const UttTables& ut, const int16_t* pcm, const double* win,
                      double* frames);
// one frame given as doubles (per-hop API)
void launch_frame_one(snmfnat_ctx* ctx, const StftGeom& g, const double* y, const double* win, double* frame);
// Y -> Ym = |Y|^pow, DC zeroed, + floor (bnmf_sep_event_RT_IS16.m:71-78); optional phase angle(Y)
void launch_stft_post(snmfnat_ctx* ctx, const StftGeom& g, const double2* Y, long long nf, double* Ym, double* Yp);
// Z = Xt^(1/pow) .* Y/|Y| with DC bins zeroed, in place on Y (synth_ifft_buff.m:10-18)
void launch_istft_pre(snmfnat_ctx* ctx, const StftGeom& g, double2* Y, const double* Xt, long long nf);
// frames <- frames/fftlen .* win (first sz samples), de-emphasis when preemph != 0 (synth_ifft_buff.m:20-26)
void launch_synth_window(snmfnat_ctx* ctx, const StftGeom& g, double* frames, const double* win, long long nf);
// overlap-add of the frames with l > delay, * overlapscale, int16 conversion (filewise_run_IS16.m:146,162-165)
void launch_ola_int16(snmfnat_ctx* ctx, const StftGeom& g, const UttTables& ut, const double* frames,
                      const double* win, int windowed, int16_t* out);

// cuFFT plan pair for nf frames
struct FftPlans {
  cufftHandle fwd = 0, inv = 0;
  long long nf = 0;
  int fftlen = 0;
  bool ok = false;
  void create(snmfnat_ctx* ctx, int fftlen, long long nf);
  void destroy();
  ~FftPlans() { destroy(); }
};

}  // namespace snmfnat
