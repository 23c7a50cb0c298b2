// MEX gateway: [x_hat_i, d_hat_i, x_tilde, g] = bnmf_sep_event_RT_IS16(y, l, g, p)
// replaces src/bnmf_sep_event_RT_IS16.m:1-423.  The struct g produced by the reference's own init_buff.m is accepted
// unchanged: on the first hop its fields seed a device-resident stream whose handle is stored in g.snmfnat_handle;
// g.B_DFT_d, the field the callers read back (src/NTF_sep_event_RT.m:137-138, SE_GUI.m), is refreshed on every hop, the
// rest stays on the device (snmfnat_stream_get reads it on demand).  A g that arrives with l == 1 starts a new file:
// the stream it carried (if any) is destroyed first, so a corpus loop does not accumulate device state.
#include <map>
#include "snmfnat_mex.h"
using namespace snmex;

static std::map<uint64_t, snmfnat_stream*>& streams() {
  static std::map<uint64_t, snmfnat_stream*> m;
  return m;
}
static uint64_t g_next_id = 1;
static void cleanup() {
  for (auto& kv : streams()) snmfnat_stream_destroy(kv.second);
  streams().clear();
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 4 || nlhs > 4) mexErrMsgIdAndTxt("snmfnat:usage", "[x_hat_i,d_hat_i,x_tilde,g] = bnmf_sep_event_RT_IS16(y,l,g,p)");
  const mxArray *y = prhs[0], *gin = prhs[2], *p = prhs[3];
  const int l = (int)mxGetScalar(prhs[1]);
  const snmfnat_params q = params(p);
  const size_t sz = q.framelength, R = q.R_x + q.R_d;
  if (mxGetNumberOfElements(y) != sz) mexErrMsgIdAndTxt("snmfnat:shape", "y must hold p.framelength samples (ch = 1)");
  mxArray* g = mxDuplicateArray(gin);          // value semantics: never write prhs
  snmfnat_stream* s = nullptr;
  const mxArray* hf = field(g, "snmfnat_handle");
  if (hf) {
    auto it = streams().find(*(uint64_t*)mxGetData(hf));
    if (it != streams().end()) {
      if (l > 1) {
        s = it->second;
      } else {   // the caller re-used a struct of a finished file: release its device state
        snmfnat_stream_destroy(it->second);
        streams().erase(it);
      }
    }
  }
  if (!s) {  // first hop of a stream: build the device state from the reference-made g (src/init_buff.m:17-62)
    const mxArray *Bmx = field(g, "B_Mel_x"), *Bmd = field(g, "B_Mel_d"), *Bx = field(g, "B_DFT_x"), *Bd = field(g, "B_DFT_d");
    const mxArray* Ad = field(g, "Ad_blk");
    if (!Bmx || !Bmd || !Bx || !Bd || !Ad) mexErrMsgIdAndTxt("snmfnat:param", "g lacks the fields init_buff creates");
    const mxArray *ws = field(p, "win_STFT"), *wi = field(p, "win_ISTFT");
    if (!ws || !wi) mexErrMsgIdAndTxt("snmfnat:param", "p.win_STFT / p.win_ISTFT missing");
    const int n1 = (int)mxGetM(Bmd), n2 = (int)mxGetM(Bx);
    check(snmfnat_stream_create(ctx(), &q, mxGetPr(ws), mxGetPr(wi), mxGetPr(Bmx), mxGetPr(Bmd), n1, mxGetPr(Bx),
                                mxGetPr(Bd), n2, mat(Ad, q.R_a, q.m_a, "g.Ad_blk"), nullptr, &s));
    if (streams().empty()) mexAtExit(cleanup);
    const uint64_t id = g_next_id++;
    streams()[id] = s;
    if (mxGetFieldNumber(g, "snmfnat_handle") < 0) mxAddField(g, "snmfnat_handle");
    mxArray* h = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t*)mxGetData(h) = id;
    mxSetField(g, 0, "snmfnat_handle", h);
  }
  seed_rng(p);                                  // sparse_nmf.m:112-114: the same H init on every hop
  mxArray* h0 = host_rand(R, 1);
  const mwSize dx[3] = {(mwSize)q.EVENT_NUM, 1, sz}, dd[3] = {1, (mwSize)q.NOISE_NUM, sz};
  mxArray* xh = (nlhs >= 1) ? mxCreateNumericArray(3, dx, mxDOUBLE_CLASS, mxREAL) : nullptr;
  mxArray* dh = (nlhs >= 2) ? mxCreateNumericArray(3, dd, mxDOUBLE_CLASS, mxREAL) : nullptr;
  mxArray* xt = mxCreateDoubleMatrix(1, sz, mxREAL);
  // the library returns class-major rows (class, sample); MATLAB wants (class,1,sample) / (1,class,sample): same
  // memory order only for one class, so go through a temporary when there are several
  std::vector<double> tx((size_t)q.EVENT_NUM * sz), td((size_t)q.NOISE_NUM * sz);
  const bool aux = nlhs >= 2 || (nlhs == 1);
  check(snmfnat_stream_step(s, mxGetPr(y), l, mxGetPr(h0), mxGetPr(xt), aux ? tx.data() : nullptr,
                            aux ? td.data() : nullptr));
  mxDestroyArray(h0);
  if (xh) for (int c = 0; c < q.EVENT_NUM; ++c) for (size_t i = 0; i < sz; ++i) mxGetPr(xh)[c + (size_t)q.EVENT_NUM * i] = tx[c * sz + i];
  if (dh) for (int c = 0; c < q.NOISE_NUM; ++c) for (size_t i = 0; i < sz; ++i) mxGetPr(dh)[c + (size_t)q.NOISE_NUM * i] = td[c * sz + i];
  if (nlhs >= 1) plhs[0] = xh;
  if (nlhs >= 2) plhs[1] = dh;
  if (nlhs >= 3) plhs[2] = xt; else mxDestroyArray(xt);
  if (nlhs >= 4) {
    // fields the callers read back: the adapted noise dictionary (src/NTF_sep_event_RT.m:137-138)
    mxArray* Bd = mxCreateDoubleMatrix(q.fftlength / 2 + 1, q.R_d, mxREAL);
    check(snmfnat_stream_get(s, "B_DFT_d", mxGetPr(Bd), (int64_t)mxGetNumberOfElements(Bd)));
    mxSetField(g, 0, "B_DFT_d", Bd);
    plhs[3] = g;
  } else {
    mxDestroyArray(g);
  }
}
