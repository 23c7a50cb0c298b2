// MEX gateway: M = mel_matrix(fs, NbCh, Nfft [, warp [, fhigh]])      replaces src/mel_matrix.m:1-40
// (dense instead of sparse; the second and third outputs of the reference have no caller and are not provided).
#include "snmfnat_mex.h"
using namespace snmex;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 3 || nlhs > 1) mexErrMsgIdAndTxt("snmfnat:usage", "M = mel_matrix(fs, NbCh, Nfft [, warp [, fhigh]])");
  const int fs = (int)mxGetScalar(prhs[0]), nb = (int)mxGetScalar(prhs[1]), nfft = (int)mxGetScalar(prhs[2]);
  const double warp = nrhs > 3 ? mxGetScalar(prhs[3]) : 1.0;
  const double fhigh = nrhs > 4 ? mxGetScalar(prhs[4]) : -1.0;
  plhs[0] = mxCreateDoubleMatrix(nfft / 2 + 1, nb, mxREAL);
  check(snmfnat_mel_matrix(fs, nb, nfft, warp, fhigh, mxGetPr(plhs[0])));
}
