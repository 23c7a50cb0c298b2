// MEX gateway: snmfnat_enhance_files(paths_in, paths_out, B_DFT_x, B_DFT_d, p [, chain_id [, B_Mel_x, B_Mel_d]])
// The hop loops of filewise_run_IS16.m:54-186 / src/NTF_sep_event_RT.m:12-156 for a whole list of files in ONE device
// call (no per-hop PCIe crossing): reads the int16 samples after the 44-byte WAV header (:92-97), enhances them with
// snmfnat_enhance_batch, writes 16-bit mono WAV files.  chain_id (optional, one per file) reproduces the B_D_u.mat
// carry-over of run_ntf_sep_RT: files with equal id are processed in list order on one adapted noise dictionary.
// With p.B_sep_mode = 'Mel' the Mel dictionaries (B_Mel_sub of the basis files) are the 7th and 8th argument.
#include <cstdio>
#include "snmfnat_mex.h"
using namespace snmex;

static std::vector<int16_t> read_pcm(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) mexErrMsgIdAndTxt("snmfnat:io", "cannot open %s", path.c_str());
  std::fseek(f, 0, SEEK_END);
  long bytes = std::ftell(f);
  std::fseek(f, 44, SEEK_SET);                       // 22 int16 of header (filewise_run_IS16.m:93-96)
  std::vector<int16_t> v(bytes > 44 ? (bytes - 44) / 2 : 0);
  if (!v.empty() && std::fread(v.data(), 2, v.size(), f) != v.size()) mexErrMsgIdAndTxt("snmfnat:io", "short read on %s", path.c_str());
  std::fclose(f);
  return v;
}
// src/pcm2wav.m:9-10: out./32767 through wavwrite(..., 16, ...) = round(x * 32768), half away from zero, clipped
static std::vector<int16_t> pcm2wav_samples(const std::vector<int16_t>& v) {
  std::vector<int16_t> y(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    const double x = (double)v[i] / 32767.0 * 32768.0;
    double r = x < 0 ? -std::floor(-x + 0.5) : std::floor(x + 0.5);
    r = r > 32767.0 ? 32767.0 : (r < -32768.0 ? -32768.0 : r);
    y[i] = (int16_t)r;
  }
  return y;
}
static void write_wav(const std::string& path, const std::vector<int16_t>& raw, int fs) {
  const std::vector<int16_t> x = pcm2wav_samples(raw);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) mexErrMsgIdAndTxt("snmfnat:io", "cannot create %s", path.c_str());
  const uint32_t data = (uint32_t)x.size() * 2, riff = 36 + data, fmt = 16, rate = fs, brate = fs * 2;
  const uint16_t pcm = 1, ch = 1, align = 2, bits = 16;
  std::fwrite("RIFF", 1, 4, f); std::fwrite(&riff, 4, 1, f); std::fwrite("WAVEfmt ", 1, 8, f); std::fwrite(&fmt, 4, 1, f);
  std::fwrite(&pcm, 2, 1, f); std::fwrite(&ch, 2, 1, f); std::fwrite(&rate, 4, 1, f); std::fwrite(&brate, 4, 1, f);
  std::fwrite(&align, 2, 1, f); std::fwrite(&bits, 2, 1, f); std::fwrite("data", 1, 4, f); std::fwrite(&data, 4, 1, f);
  std::fwrite(x.data(), 2, x.size(), f);
  std::fclose(f);
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  (void)plhs;
  if (nrhs < 5 || nlhs > 0)
    mexErrMsgIdAndTxt("snmfnat:usage", "snmfnat_enhance_files(paths_in,paths_out,B_DFT_x,B_DFT_d,p[,chain_id[,B_Mel_x,B_Mel_d]])");
  if (!mxIsCell(prhs[0]) || !mxIsCell(prhs[1])) mexErrMsgIdAndTxt("snmfnat:usage", "paths must be cell arrays of strings");
  const size_t n = mxGetNumberOfElements(prhs[0]);
  const mxArray* p = prhs[4];
  const snmfnat_params q = params(p);
  const size_t F = q.fftlength / 2 + 1, R = q.R_x + q.R_d;
  std::vector<std::vector<int16_t>> pcm(n), out(n);
  std::vector<const int16_t*> pin(n);
  std::vector<int16_t*> pout(n);
  std::vector<int64_t> len(n);
  std::vector<std::string> po(n);
  for (size_t i = 0; i < n; ++i) {
    char* a = mxArrayToString(mxGetCell(prhs[0], i));
    char* b = mxArrayToString(mxGetCell(prhs[1], i));
    pcm[i] = read_pcm(a);
    po[i] = b;
    mxFree(a); mxFree(b);
    len[i] = (int64_t)pcm[i].size();
    out[i].resize((size_t)(len[i] / q.frameshift + 1) * q.frameshift);
    pin[i] = pcm[i].data();
    pout[i] = out[i].data();
  }
  std::vector<int32_t> chain;
  if (nrhs > 5 && !mxIsEmpty(prhs[5])) for (size_t i = 0; i < n; ++i) chain.push_back((int32_t)mxGetPr(prhs[5])[i]);
  // RNG on the host: H init after rand('seed',.), then per file the two draws of init_buff.m:37-38 in their order:
  // g.A_d = rand(R_d, m) (overwritten before its first use, but it advances the generator) and g.Ad_blk = rand(R_a, m_a)
  seed_rng(p);
  mxArray* h0 = host_rand(R, 1);
  std::vector<double> ad((size_t)n * q.R_a * q.m_a);
  for (size_t i = 0; i < n; ++i) {
    mxDestroyArray(host_rand(q.R_d, 1));
    mxArray* a = host_rand(q.R_a, q.m_a);
    std::memcpy(ad.data() + i * q.R_a * q.m_a, mxGetPr(a), sizeof(double) * q.R_a * q.m_a);
    mxDestroyArray(a);
  }
  const mxArray *ws = field(p, "win_STFT"), *wi = field(p, "win_ISTFT");
  if (!ws || !wi) mexErrMsgIdAndTxt("snmfnat:param", "p.win_STFT / p.win_ISTFT missing");
  // p.gpus = [0 1 2 ...] (optional): the files are split over these devices, one host thread per device
  // (Do_MultiBatch_IS16_20160324_CHiME4.m:202-208 walks the corpus file by file; the files are independent)
  const mxArray* gp = field(p, "gpus");
  if (gp && mxGetNumberOfElements(gp) > 1 && q.B_sep_mode != SNMFNAT_SEP_MEL) {
    std::vector<int> dev;
    for (size_t i = 0; i < mxGetNumberOfElements(gp); ++i) dev.push_back((int)mxGetPr(gp)[i]);
    const int rcm = snmfnat_enhance_batch_multi(dev.data(), (int)dev.size(), &q, mxGetPr(ws), mxGetPr(wi),
                                                mat(prhs[2], F, q.R_x, "B_DFT_x"), mat(prhs[3], F, q.R_d, "B_DFT_d"), (int)F,
                                                (int)n, pin.data(), len.data(), chain.empty() ? nullptr : chain.data(),
                                                mxGetPr(h0), ad.data(), (int64_t)q.R_a * q.m_a, pout.data());
    mxDestroyArray(h0);
    check(rcm);
    for (size_t i = 0; i < n; ++i) write_wav(po[i], out[i], q.fs);
    return;
  }
  snmfnat_batch* bt = nullptr;
  check(snmfnat_batch_create(ctx(), &q, mxGetPr(ws), mxGetPr(wi), mat(prhs[2], F, q.R_x, "B_DFT_x"),
                             mat(prhs[3], F, q.R_d, "B_DFT_d"), (int)F, (int)n, len.data(),
                             chain.empty() ? nullptr : chain.data(), mxGetPr(h0), ad.data(), (int64_t)q.R_a * q.m_a, &bt));
  int rc = 0;
  if (q.B_sep_mode == SNMFNAT_SEP_MEL) {   // filewise_run_IS16.m:46-51: B1_x / B1_d are the Mel dictionaries
    if (nrhs < 8) {
      snmfnat_batch_destroy(bt);
      mexErrMsgIdAndTxt("snmfnat:usage", "p.B_sep_mode = 'Mel' needs B_Mel_x and B_Mel_d (arguments 7 and 8)");
    }
    const size_t n1 = mxGetM(prhs[6]);
    rc = snmfnat_batch_set_mel(bt, mat(prhs[6], n1, q.R_x, "B_Mel_x"), mat(prhs[7], n1, q.R_d, "B_Mel_d"), (int)n1, nullptr);
  }
  if (!rc) rc = snmfnat_batch_upload(bt, pin.data());
  if (!rc) rc = snmfnat_batch_run(bt);
  if (!rc) rc = snmfnat_batch_download(bt, pout.data());
  snmfnat_batch_destroy(bt);
  check(rc);
  mxDestroyArray(h0);
  for (size_t i = 0; i < n; ++i) write_wav(po[i], out[i], q.fs);
}
