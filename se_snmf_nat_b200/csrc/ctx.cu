// Context, error plumbing and the default parameter set of libsnmfnat.
#include <cstring>
#include <cmath>
#include "common.cuh"

namespace snmfnat {

static thread_local std::string g_last_error;

std::string vformat(const char* fmt, va_list ap) {
  char buf[1024];
  vsnprintf(buf, sizeof(buf), fmt, ap);
  return std::string(buf);
}
void fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::string m = vformat(fmt, ap);
  va_end(ap);
  throw Error(code, m);
}
void set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }

void check_launch(snmfnat_ctx* ctx, const char* what) {
  (void)ctx;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(SNMFNAT_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
}

}  // namespace snmfnat

using namespace snmfnat;

extern "C" {

int snmfnat_version(void) { return SNMFNAT_VERSION; }

const char* snmfnat_last_error(const snmfnat_ctx* ctx) {
  (void)ctx;
  return g_last_error.c_str();
}

int snmfnat_ctx_create(int device, snmfnat_ctx** out) {
  SN_API_BEGIN
  SN_REQUIRE(out != nullptr, SNMFNAT_EINVAL, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    fail(SNMFNAT_ENODEVICE, "no CUDA device available (%s); libsnmfnat has no CPU fallback",
         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  SN_REQUIRE(device >= 0 && device < count, SNMFNAT_ENODEVICE, "device %d out of range (have %d)", device, count);
  cudaDeviceProp prop;
  SN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    fail(SNMFNAT_ENODEVICE, "device %d is sm_%d%d; libsnmfnat is built for sm_100a (B200) only", device, prop.major,
         prop.minor);
  SN_CUDA(cudaSetDevice(device));
  snmfnat_ctx* c = new snmfnat_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    fail(SNMFNAT_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
  }
  *out = c;
  SN_API_END
}

int snmfnat_ctx_destroy(snmfnat_ctx* ctx) {
  SN_API_BEGIN
  if (ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
  }
  SN_API_END
}

int snmfnat_ctx_sync(snmfnat_ctx* ctx) {
  SN_API_BEGIN
  SN_REQUIRE(ctx != nullptr, SNMFNAT_EINVAL, "ctx is NULL");
  SN_CUDA(cudaStreamSynchronize(ctx->stream));
  SN_API_END
}

void* snmfnat_ctx_cuda_stream(snmfnat_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int64_t snmfnat_ctx_launch_count(const snmfnat_ctx* ctx) { return ctx ? ctx->launches : 0; }

// settings/initial_setting_SNMF_NAT.m:1-149
void snmfnat_params_default(snmfnat_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->fs = 16000;
  p->framelength = 640;   // round(0.040*fs)      :25
  p->frameshift = 160;    // round(0.010*fs)      :26
  p->fftlength = 1024;    // 2^ceil(log2(640))    :29
  p->blk_len_sep = 1; p->blk_hop_sep = 1; p->Splice = 0;           // :16-18
  p->delay = 0 + 1 + (int)std::floor(0.040 / 0.010 / 2 + 0.5);     // :27  -> 3
  p->EVENT_NUM = 1; p->NOISE_NUM = 1; p->EVENT_RANK[0] = 1; p->NOISE_RANK[0] = 1;  // :40-44
  p->R_x = 100; p->R_d = 100;                                      // :48-49
  p->R_a = 50; p->m_a = 100; p->init_N_len = 15; p->adapt_train_N = 1;  // :56-59
  p->blk_sparse = 1; p->P_len_k = 60; p->P_len_l = 20; p->blk_gap = 3;  // :64-70
  p->DCbin = 5; p->DCbin_back = 5; p->F_order = 64;                // :31,90-92
  p->B_sep_mode = SNMFNAT_SEP_DFT; p->MelConv = 1;                 // :98-99
  p->cf = SNMFNAT_CF_KL; p->max_iter = 100; p->cost_check = 1;     // :106-112
  p->basis_update_N = 0; p->basis_update_E = 0;                    // :113-114
  p->ENHANCE_METHOD = SNMFNAT_ENH_MMSE;                            // :117
  p->overlapscale = 2.0 * 160 / 640;                               // :32
  p->pow = 2.0; p->nonzerofloor = 1e-9;                            // :37,53
  p->overlap_m_a = 0.01; p->Ar_up = 1.0;                           // :60-61
  p->alpha_p = 0.4; p->preemph = 0.0;                              // :69,88
  p->beta_div = 1.0;
  p->sparsity = 5.0; p->conv_eps = 1e-3;                           // :107,109
  p->alpha_eta = 0.4; p->alpha_d = 0.6; p->beta = 1.0; p->beta_max = 1000.0;  // :119,133,138-139
  p->sparsity_mdi = 5.0; p->conv_eps_mdi = 1e-5;                   // :75-76
}

}  // extern "C"
