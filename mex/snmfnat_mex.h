// Shared helpers of the MEX gateways: marshalling of the reference's `p` struct and matrices to the C ABI of
// libsnmfnat (include/snmfnat.h).  The gateways hold no numerics: every computation happens in the CUDA library.
#pragma once
#ifdef SNMFNAT_MEX_SHIM
#include "mex_shim.h"
#else
#include "mex.h"
#endif
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../include/snmfnat.h"

namespace snmex {

inline snmfnat_ctx*& ctx_slot() {
  static snmfnat_ctx* c = nullptr;
  return c;
}
inline void at_exit() {
  if (ctx_slot()) snmfnat_ctx_destroy(ctx_slot());
  ctx_slot() = nullptr;
}
// One context per MATLAB process (device = $SNMFNAT_DEVICE, default 0); no CPU fallback: a missing device is an error.
// Multi-GPU corpus runs go through snmfnat_enhance_batch_multi, which creates one context per device itself.
inline snmfnat_ctx* ctx() {
  if (!ctx_slot()) {
    const char* e = std::getenv("SNMFNAT_DEVICE");
    if (snmfnat_ctx_create(e ? std::atoi(e) : 0, &ctx_slot()) != 0) mexErrMsgIdAndTxt("snmfnat:device", "%s", snmfnat_last_error(nullptr));
    mexLock();
    mexAtExit(at_exit);
  }
  return ctx_slot();
}
inline void check(int rc) {
  if (rc != 0) mexErrMsgIdAndTxt("snmfnat:error", "%s", snmfnat_last_error(nullptr));
}
inline const mxArray* field(const mxArray* s, const char* name) { return mxIsStruct(s) ? mxGetField(s, 0, name) : nullptr; }
inline bool has(const mxArray* s, const char* name) { return field(s, name) != nullptr; }
inline double num(const mxArray* s, const char* name, double dflt) {
  const mxArray* f = field(s, name);
  return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}
inline const double* mat(const mxArray* a, size_t rows, size_t cols, const char* what) {
  if (!a || !mxIsDouble(a) || mxGetM(a) != rows || mxGetN(a) != cols)
    mexErrMsgIdAndTxt("snmfnat:shape", "%s must be a %d x %d double matrix", what, (int)rows, (int)cols);
  return mxGetPr(a);
}
inline std::string str(const mxArray* s, const char* name, const char* dflt) {
  const mxArray* f = field(s, name);
  if (!f || !mxIsChar(f)) return dflt;
  char* c = mxArrayToString(f);
  std::string r(c);
  mxFree(c);
  return r;
}
// logical or numeric index vector -> uint8
inline std::vector<uint8_t> logicals(const mxArray* a, size_t r) {
  std::vector<uint8_t> v(r, 1);
  if (!a) return v;
  if (mxGetNumberOfElements(a) != r) mexErrMsgIdAndTxt("snmfnat:shape", "update index vector must have r entries");
  if (mxIsLogical(a)) {
    const mxLogical* l = mxGetLogicals(a);
    for (size_t i = 0; i < r; ++i) v[i] = l[i] ? 1 : 0;
  } else {
    const double* d = mxGetPr(a);
    for (size_t i = 0; i < r; ++i) v[i] = d[i] != 0;
  }
  return v;
}

// global p (settings/initial_setting_SNMF_NAT.m) -> snmfnat_params.  Missing fields keep the shipped defaults.
inline snmfnat_params params(const mxArray* p) {
  snmfnat_params q;
  snmfnat_params_default(&q);
#define SN_I(name) q.name = (int32_t)num(p, #name, q.name)
#define SN_D(name) q.name = num(p, #name, q.name)
  SN_I(fs); SN_I(framelength); SN_I(frameshift); SN_I(fftlength); SN_I(delay);
  SN_I(blk_len_sep); SN_I(blk_hop_sep); SN_I(Splice); SN_I(EVENT_NUM); SN_I(NOISE_NUM);
  SN_I(R_x); SN_I(R_d); SN_I(R_a); SN_I(m_a); SN_I(init_N_len); SN_I(adapt_train_N);
  SN_I(blk_sparse); SN_I(P_len_k); SN_I(P_len_l); SN_I(blk_gap); SN_I(DCbin); SN_I(DCbin_back); SN_I(F_order);
  SN_I(MelConv); SN_I(max_iter); SN_I(cost_check); SN_I(basis_update_N); SN_I(basis_update_E);
  SN_D(overlapscale); SN_D(pow); SN_D(nonzerofloor); SN_D(overlap_m_a); SN_D(Ar_up); SN_D(alpha_p); SN_D(preemph);
  SN_D(sparsity); SN_D(conv_eps); SN_D(alpha_eta); SN_D(alpha_d); SN_D(beta); SN_D(beta_max);
  SN_D(sparsity_mdi); SN_D(conv_eps_mdi);
#undef SN_I
#undef SN_D
  const char* rk[2] = {"EVENT_RANK", "NOISE_RANK"};
  for (int w = 0; w < 2; ++w) {
    const mxArray* f = field(p, rk[w]);
    if (!f) continue;
    const size_t n = mxGetNumberOfElements(f);
    if (n > SNMFNAT_MAX_CLASSES) mexErrMsgIdAndTxt("snmfnat:shape", "%s has too many classes", rk[w]);
    for (size_t i = 0; i < n; ++i) (w == 0 ? q.EVENT_RANK : q.NOISE_RANK)[i] = (int32_t)mxGetPr(f)[i];
  }
  const std::string cf = str(p, "cf", "kl");
  q.cf = cf == "is" ? SNMFNAT_CF_IS : cf == "kl" ? SNMFNAT_CF_KL : cf == "ed" ? SNMFNAT_CF_ED : SNMFNAT_CF_BETA;
  q.beta_div = num(p, "beta", 1.0);   // only read when cf is none of is/kl/ed (sparse_nmf.m:106-109)
  q.ENHANCE_METHOD = str(p, "ENHANCE_METHOD", "MMSE") == "Wiener" ? SNMFNAT_ENH_WIENER : SNMFNAT_ENH_MMSE;
  q.B_sep_mode = str(p, "B_sep_mode", "DFT") == "Mel" ? SNMFNAT_SEP_MEL : SNMFNAT_SEP_DFT;
  return q;
}

// the optional fields of p read by sparse_nmf (sparse_nmf.m:79-164,260)
inline snmfnat_nmf_opts nmf_opts(const mxArray* p, const char* sparsity_field, const char* eps_field,
                                 std::vector<double>& sparsity, size_t r, size_t n) {
  snmfnat_nmf_opts o;
  std::memset(&o, 0, sizeof(o));
  o.max_iter = (int32_t)num(p, "max_iter", 100);
  const std::string cf = str(p, "cf", "kl");
  o.cf = cf == "is" ? SNMFNAT_CF_IS : cf == "kl" ? SNMFNAT_CF_KL : cf == "ed" ? SNMFNAT_CF_ED : SNMFNAT_CF_BETA;
  o.beta_div = num(p, "beta", 1.0);
  if (!has(p, "cost_check")) mexErrMsgIdAndTxt("snmfnat:param", "p.cost_check is required (sparse_nmf.m:260)");
  o.cost_check = num(p, "cost_check", 1) != 0;
  o.conv_eps = num(p, eps_field, 0.0);
  const mxArray* s = field(p, sparsity_field);
  if (!s) {
    sparsity.assign(1, 0.0);
    o.sparsity_rows = o.sparsity_cols = 1;
  } else {
    const size_t m = mxGetM(s), c = mxGetN(s), ne = mxGetNumberOfElements(s);
    sparsity.assign(mxGetPr(s), mxGetPr(s) + ne);
    if (ne == 1) { o.sparsity_rows = o.sparsity_cols = 1; }
    else if (c == 1 && m == r) { o.sparsity_rows = (int32_t)r; o.sparsity_cols = 1; }
    else if (m == r && c == n) { o.sparsity_rows = (int32_t)r; o.sparsity_cols = (int32_t)n; }
    else mexErrMsgIdAndTxt("snmfnat:shape", "p.%s must be scalar, r x 1 or r x n", sparsity_field);
  }
  return o;
}

// rand('seed', p.random_seed); then rand(m, n) on the HOST interpreter, so that host and device see MATLAB's numbers
inline void seed_rng(const mxArray* p) {
  const double seed = num(p, "random_seed", 1);
  if (seed > 0) {   // sparse_nmf.m:112-114
    mxArray* in[2] = {mxCreateString("seed"), mxCreateDoubleScalar(seed)};
    mexCallMATLAB(0, nullptr, 2, in, "rand");
    mxDestroyArray(in[0]);
    mxDestroyArray(in[1]);
  }
}
inline mxArray* host_rand(size_t m, size_t n) {
  mxArray* in[2] = {mxCreateDoubleScalar((double)m), mxCreateDoubleScalar((double)n)};
  mxArray* out = nullptr;
  mexCallMATLAB(1, &out, 2, in, "rand");
  mxDestroyArray(in[0]);
  mxDestroyArray(in[1]);
  return out;
}

}  // namespace snmex
