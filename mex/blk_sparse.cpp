// MEX gateway: [Q, r_blk_out] = blk_sparse(X, D, r_blk, l, p)      replaces src/blk_sparse.m:1-37
#include "snmfnat_mex.h"
using namespace snmex;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 5 || nlhs > 2) mexErrMsgIdAndTxt("snmfnat:usage", "[Q,r_blk_out] = blk_sparse(X,D,r_blk,l,p)");
  const size_t K = mxGetM(prhs[0]);
  const snmfnat_params q = params(prhs[4]);
  plhs[0] = mxCreateDoubleMatrix(K, 1, mxREAL);
  mxArray* ro = mxCreateDoubleMatrix(K, q.P_len_l, mxREAL);
  check(snmfnat_blk_sparse(ctx(), mat(prhs[0], K, 1, "X"), mat(prhs[1], K, 1, "D"), mat(prhs[2], K, q.P_len_l, "r_blk"),
                           (int)K, (int)mxGetScalar(prhs[3]), &q, mxGetPr(plhs[0]), mxGetPr(ro)));
  if (nlhs > 1) plhs[1] = ro; else mxDestroyArray(ro);
}
