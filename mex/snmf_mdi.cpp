// MEX gateway: [v_MDI, h, objective] = snmf_mdi(v, Dm, p)      replaces src/snmf_mdi.m
// Build a second copy with -DSNMFNAT_SOFT_MASK as snmf_mdi_Sm (src/snmf_mdi_Sm.m).
#include "snmfnat_mex.h"
using namespace snmex;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 3 || nlhs > 3) mexErrMsgIdAndTxt("snmfnat:usage", "[v_MDI,h,objective] = snmf_mdi(v,Dm,p)");
  const mxArray *v = prhs[0], *Dm = prhs[1], *p = prhs[2];
  const size_t m = mxGetM(v), n = mxGetN(v);
  seed_rng(p);
  const mxArray* iw = field(p, "init_w");
  if (!iw) mexErrMsgIdAndTxt("snmfnat:param", "p.init_w is required");
  const size_t r = mxGetN(iw);
  std::vector<double> h0, mk(m * n);
  const mxArray* ih = field(p, "init_h");
  if (!ih) {
    mxArray* rh = host_rand(r, n);
    h0.assign(mxGetPr(rh), mxGetPr(rh) + r * n);
    mxDestroyArray(rh);
  } else {
    h0.assign(mat(ih, r, n, "p.init_h"), mat(ih, r, n, "p.init_h") + r * n);
  }
  if (mxIsLogical(Dm)) for (size_t i = 0; i < m * n; ++i) mk[i] = mxGetLogicals(Dm)[i] ? 1.0 : 0.0;
  else std::memcpy(mk.data(), mat(Dm, m, n, "Dm"), m * n * sizeof(double));
  if (!has(p, "sparsity_mdi") || !has(p, "conv_eps_mdi"))   // the defaults at snmf_mdi.m:93-99 test the wrong names
    mexErrMsgIdAndTxt("snmfnat:param", "p.sparsity_mdi and p.conv_eps_mdi are required");
  std::vector<double> sp;
  snmfnat_nmf_opts o = nmf_opts(p, "sparsity_mdi", "conv_eps_mdi", sp, r, n);
  std::vector<uint8_t> wi = logicals(field(p, "w_update_ind"), r), hi = logicals(field(p, "h_update_ind"), r);
  plhs[0] = mxCreateDoubleMatrix(m, n, mxREAL);
  mxArray* h = mxCreateDoubleMatrix(r, n, mxREAL);
  std::vector<double> div(o.max_iter > 0 ? o.max_iter : 1), cost(div.size());
  int its = 0;
#ifdef SNMFNAT_SOFT_MASK
  const int soft = 1;
#else
  const int soft = 0;
#endif
  check(snmfnat_snmf_mdi(ctx(), mat(v, m, n, "v"), mk.data(), soft, (int)m, (int)n, (int)r, &o, sp.data(),
                         mat(iw, m, r, "p.init_w"), h0.data(), wi.data(), hi.data(), mxGetPr(plhs[0]), mxGetPr(h),
                         div.data(), cost.data(), &its));
  if (nlhs > 1) plhs[1] = h; else mxDestroyArray(h);
  if (nlhs > 2) {
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
