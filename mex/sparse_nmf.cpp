// MEX gateway: [w, h, objective] = sparse_nmf(v, p)        replaces src/sparse_nmf.m:1-292
#include "snmfnat_mex.h"
using namespace snmex;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 2 || nlhs > 3) mexErrMsgIdAndTxt("snmfnat:usage", "[w,h,objective] = sparse_nmf(v,p)");
  const mxArray* v = prhs[0];
  const mxArray* p = prhs[1];
  const size_t m = mxGetM(v), n = mxGetN(v);
  seed_rng(p);                                                           // :112-114
  // init_w / init_h / r exactly as :116-140, random parts drawn by the host's rand
  std::vector<double> w0, h0;
  size_t r;
  const mxArray* iw = field(p, "init_w");
  if (!iw) {
    if (!has(p, "r")) mexErrMsgIdAndTxt("snmfnat:param", "Number of components or initialization must be given");
    r = (size_t)num(p, "r", 0);
    mxArray* rw = host_rand(m, r);
    w0.assign(mxGetPr(rw), mxGetPr(rw) + m * r);
    mxDestroyArray(rw);
  } else {
    const size_t ri = mxGetN(iw);
    w0.assign(mxGetPr(iw), mxGetPr(iw) + m * ri);
    r = ri;
    if (has(p, "r") && ri < (size_t)num(p, "r", 0)) {
      r = (size_t)num(p, "r", 0);
      mxArray* rw = host_rand(m, r - ri);
      w0.insert(w0.end(), mxGetPr(rw), mxGetPr(rw) + m * (r - ri));
      mxDestroyArray(rw);
    }
  }
  const mxArray* ih = field(p, "init_h");
  if (!ih) {
    mxArray* rh = host_rand(r, n);
    h0.assign(mxGetPr(rh), mxGetPr(rh) + r * n);
    mxDestroyArray(rh);
  } else if (mxIsChar(ih)) {
    h0.assign(r * n, 1.0);                                               // 'ones' (:135-137)
  } else {
    h0.assign(mat(ih, r, n, "p.init_h"), mat(ih, r, n, "p.init_h") + r * n);
  }
  std::vector<double> sp;
  snmfnat_nmf_opts o = nmf_opts(p, "sparsity", "conv_eps", sp, r, n);
  std::vector<uint8_t> wi = logicals(field(p, "w_update_ind"), r), hi = logicals(field(p, "h_update_ind"), r);
  plhs[0] = mxCreateDoubleMatrix(m, r, mxREAL);
  mxArray* h = mxCreateDoubleMatrix(r, n, mxREAL);
  std::vector<double> div(o.max_iter > 0 ? o.max_iter : 1), cost(div.size());
  int its = 0;
  // Dictionary training (run_basis_train.m:84-88: W and H both updated, KL, a whole corpus of frames) goes to the
  // tensor-core path (snmfnat_train_*: tf32 operands, fp32 state; W/H within 1e-3 of the float64 path) when the caller
  // opts in with p.useGPU ~= 0 -- the reference's own switch between sparse_nmf and sparse_nmf_GPU
  // (bnmf_sep_event_RT_IS16.m:150-154) -- and the problem has its shape: every atom updated, r <= 256, n >= 16384,
  // scalar sparsity.  Everything else runs the float64 kernels.
  bool all_upd = true;
  for (size_t i = 0; i < r; ++i) all_upd = all_upd && wi[i] && hi[i];
  const bool tensor_path = all_upd && o.cf == SNMFNAT_CF_KL && r <= 256 && n >= 16384 && sp.size() == 1 &&
                           num(p, "useGPU", 0) != 0 && o.max_iter > 0;
  if (tensor_path) {
    const double* vd = mat(v, m, n, "v");
    std::vector<float> vf(m * n), wf(w0.begin(), w0.end()), hf(h0.begin(), h0.end());
    for (size_t i = 0; i < m * n; ++i) vf[i] = (float)vd[i];
    snmfnat_train* t = nullptr;
    check(snmfnat_train_create(ctx(), (int)m, (int)r, (int64_t)n, sp[0], 1, &t));
    int rc = snmfnat_train_set_data(t, vf.data(), 0, wf.data(), hf.data(), 0);
    if (!rc) {
      if (o.cost_check) {
        rc = snmfnat_train_run(t, o.max_iter, o.conv_eps, div.data(), cost.data(), &its);
      } else {                                                           // no cost evaluation, no early stop (:260)
        rc = snmfnat_train_iterate(t, o.max_iter, nullptr, nullptr);
        its = o.max_iter;
      }
    }
    if (!rc) rc = snmfnat_train_get_w(t, wf.data());
    if (!rc) rc = snmfnat_train_get_h(t, hf.data(), 0, (int64_t)n);
    snmfnat_train_destroy(t);
    check(rc);
    for (size_t i = 0; i < m * r; ++i) mxGetPr(plhs[0])[i] = wf[i];
    for (size_t i = 0; i < r * n; ++i) mxGetPr(h)[i] = hf[i];
  } else {
    check(snmfnat_sparse_nmf(ctx(), mat(v, m, n, "v"), (int)m, (int)n, (int)r, &o, sp.data(), w0.data(), h0.data(),
                             wi.data(), hi.data(), mxGetPr(plhs[0]), mxGetPr(h), div.data(), cost.data(), &its));
  }
  if (nlhs > 1) plhs[1] = h; else mxDestroyArray(h);
  if (nlhs > 2) {                                                        // objective.div / .cost (:171-173,279-280)
    const char* names[2] = {"div", "cost"};
    plhs[2] = mxCreateStructMatrix(1, 1, 2, names);
    const size_t len = (o.cost_check && its < o.max_iter) ? its : o.max_iter;
    mxArray* d = mxCreateDoubleMatrix(1, len, mxREAL);
    mxArray* c = mxCreateDoubleMatrix(1, len, mxREAL);
    std::memcpy(mxGetPr(d), div.data(), len * sizeof(double));
    std::memcpy(mxGetPr(c), cost.data(), len * sizeof(double));
    mxSetField(plhs[2], 0, "div", d);
    mxSetField(plhs[2], 0, "cost", c);
  }
}
