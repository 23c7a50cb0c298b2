// MEX gateway: [C, A] = GIST_NTF(p, B, S_mag)        replaces src/GIST_NTF.m:1-160
// (compile with -DSNMFNAT_NTF_C for GIST_NTF_C, src/GIST_NTF_C.m: the objective is evaluated only when p.cost_check)
#include "snmfnat_mex.h"
using namespace snmex;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 3 || nlhs > 2) mexErrMsgIdAndTxt("snmfnat:usage", "[C,A] = GIST_NTF(p, B, S_mag)");
  const mxArray *p = prhs[0], *B = prhs[1], *S = prhs[2];
  if (mxGetNumberOfDimensions(S) != 3) mexErrMsgIdAndTxt("snmfnat:shape", "S_mag must be Channel x N x M");
  const mwSize* dims = mxGetDimensions(S);
  const size_t Ch = dims[0], N = dims[1], M = dims[2], K = mxGetN(B);
  mxArray* c0 = host_rand(Ch, K);                     // C = rand(Channel, K)                        GIST_NTF.m:14
#ifdef SNMFNAT_NTF_C
  const int cc = num(p, "cost_check", 1) != 0 ? 1 : 0;
#else
  const int cc = -1;
#endif
  const int max_iter = (int)num(p, "max_iter", 100);
  plhs[0] = mxCreateDoubleMatrix(Ch, K, mxREAL);
  int its = 0;
  check(snmfnat_gist_ntf(ctx(), mxGetPr(S), (int)Ch, (int)N, (int)M, mat(B, N, K, "B"), (int)K, mxGetPr(c0), nullptr,
                         num(p, "sparsity", 0.0), num(p, "nonzerofloor", 1e-9), max_iter, num(p, "conv_eps", 0.0), cc,
                         mxGetPr(plhs[0]), nullptr, nullptr, &its));
  mxDestroyArray(c0);
  if (nlhs > 1) {                                     // A = ones(M, K), never updated (A_UPDATE = 0)          :6,16
    plhs[1] = mxCreateDoubleMatrix(M, K, mxREAL);
    double* a = mxGetPr(plhs[1]);
    for (size_t i = 0; i < M * K; ++i) a[i] = 1.0;
  }
}
