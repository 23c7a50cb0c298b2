// MEX gateway: [S_mag, S_phase] = stft_fft(s, sz, shift, fftlen, DCbin, win, preemph)   replaces src/stft_fft.m:1-37
#include "snmfnat_mex.h"
using namespace snmex;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 7 || nlhs > 2) mexErrMsgIdAndTxt("snmfnat:usage", "[S_mag,S_phase] = stft_fft(s,sz,shift,fftlen,DCbin,win,preemph)");
  const size_t len = mxGetNumberOfElements(prhs[0]);
  const int sz = (int)mxGetScalar(prhs[1]), shift = (int)mxGetScalar(prhs[2]), fftlen = (int)mxGetScalar(prhs[3]);
  const int DCbin = (int)mxGetScalar(prhs[4]);
  if (mxGetNumberOfElements(prhs[5]) != (size_t)sz) mexErrMsgIdAndTxt("snmfnat:shape", "win must have sz entries");
  const size_t half = fftlen / 2 + 1, nfr = len / shift;
  plhs[0] = mxCreateDoubleMatrix(half, nfr, mxREAL);
  mxArray* ph = mxCreateDoubleMatrix(half, nfr, mxREAL);
  check(snmfnat_stft_fft(ctx(), mxGetPr(prhs[0]), (int64_t)len, sz, shift, fftlen, DCbin, mxGetPr(prhs[5]),
                         mxGetScalar(prhs[6]), mxGetPr(plhs[0]), mxGetPr(ph)));
  if (nlhs > 1) plhs[1] = ph; else mxDestroyArray(ph);
}
