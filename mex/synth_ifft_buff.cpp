// MEX gateway: s_buff = synth_ifft_buff(TF_mag, TF_phase, sz, fftlen, win, preemph, DCbin_back, pow)
// replaces src/synth_ifft_buff.m:1-33
#include "snmfnat_mex.h"
using namespace snmex;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 8 || nlhs > 1) mexErrMsgIdAndTxt("snmfnat:usage", "s_buff = synth_ifft_buff(TF_mag,TF_phase,sz,fftlen,win,preemph,DCbin_back,pow)");
  const size_t fn = mxGetM(prhs[0]), frames = mxGetN(prhs[0]);
  const int sz = (int)mxGetScalar(prhs[2]), fftlen = (int)mxGetScalar(prhs[3]);
  plhs[0] = mxCreateDoubleMatrix(sz, frames, mxREAL);
  check(snmfnat_synth_ifft_buff(ctx(), mxGetPr(prhs[0]), mxGetPr(prhs[1]), (int)fn, (int)frames, sz, fftlen,
                                mxGetPr(prhs[4]), mxGetScalar(prhs[5]), (int)mxGetScalar(prhs[6]), mxGetScalar(prhs[7]),
                                mxGetPr(plhs[0])));
}
