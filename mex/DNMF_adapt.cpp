// MEX gateway: B_a = DNMF_adapt(Y, D, B, p)        replaces src/DNMF_adapt.m:1-21
#include "snmfnat_mex.h"
using namespace snmex;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 4 || nlhs > 1) mexErrMsgIdAndTxt("snmfnat:usage", "B_a = DNMF_adapt(Y,D,B,p)");
  const mxArray *Y = prhs[0], *D = prhs[1], *B = prhs[2], *p = prhs[3];
  const size_t F = mxGetM(Y), n = mxGetN(Y);
  const int R_x = (int)num(p, "R_x", 0), R_d = (int)num(p, "R_d", 0);
  const size_t r = (size_t)(R_x + R_d);
  seed_rng(p);                              // the inner H-solve draws rand(r, n) right after the reseed
  mxArray* rh = host_rand(r, n);
  std::vector<double> sp;
  snmfnat_nmf_opts o = nmf_opts(p, "sparsity", "conv_eps", sp, r, n);
  plhs[0] = mxCreateDoubleMatrix(F, R_d, mxREAL);
  check(snmfnat_dnmf_adapt(ctx(), mat(Y, F, n, "Y"), mat(D, F, n, "D"), mat(B, F, r, "B"), (int)F, (int)n, R_x, R_d, &o,
                           sp.data(), mxGetPr(rh), mxGetPr(plhs[0])));
  mxDestroyArray(rh);
}
