// Entry points declared in include/snmfnat.h whose implementation has not landed yet: they fail loudly.
#include "common.cuh"
using namespace snmfnat;
#define SN_STUB(name) { SN_API_BEGIN fail(SNMFNAT_EUNSUPPORTED, name " is not implemented yet"); SN_API_END }
extern "C" {
int snmfnat_train_create(snmfnat_ctx*, int, int, int64_t, double, int, snmfnat_train**) SN_STUB("snmfnat_train_create")
int snmfnat_train_destroy(snmfnat_train*) SN_STUB("snmfnat_train_destroy")
int snmfnat_train_nccl_unique_id(void*) SN_STUB("snmfnat_train_nccl_unique_id")
int snmfnat_train_attach_nccl(snmfnat_train*, const void*, int, int) SN_STUB("snmfnat_train_attach_nccl")
int snmfnat_train_set_data(snmfnat_train*, const float*, int, const float*, const float*, int) SN_STUB("snmfnat_train_set_data")
void* snmfnat_train_dev_ptr(snmfnat_train*, const char*) { return nullptr; }
int snmfnat_train_reset(snmfnat_train*) SN_STUB("snmfnat_train_reset")
int snmfnat_train_iterate(snmfnat_train*, int, double*, double*) SN_STUB("snmfnat_train_iterate")
int snmfnat_train_get_w(snmfnat_train*, float*) SN_STUB("snmfnat_train_get_w")
int snmfnat_train_get_h(snmfnat_train*, float*, int64_t, int64_t) SN_STUB("snmfnat_train_get_h")
int snmfnat_sparse_nmf(snmfnat_ctx*, const double*, int, int, int, const snmfnat_nmf_opts*, const double*, const double*, const double*, const uint8_t*, const uint8_t*, double*, double*, double*, double*, int*) SN_STUB("snmfnat_sparse_nmf")
int snmfnat_snmf_mdi(snmfnat_ctx*, const double*, const double*, int, int, int, int, const snmfnat_nmf_opts*, const double*, const double*, const double*, const uint8_t*, const uint8_t*, double*, double*, double*, double*, int*) SN_STUB("snmfnat_snmf_mdi")
int snmfnat_dnmf_adapt(snmfnat_ctx*, const double*, const double*, const double*, int, int, int, int, const snmfnat_nmf_opts*, const double*, const double*, double*) SN_STUB("snmfnat_dnmf_adapt")
int snmfnat_stft_fft(snmfnat_ctx*, const double*, int64_t, int, int, int, int, const double*, double, double*, double*) SN_STUB("snmfnat_stft_fft")
int snmfnat_synth_ifft_buff(snmfnat_ctx*, const double*, const double*, int, int, int, int, const double*, double, int, double, double*) SN_STUB("snmfnat_synth_ifft_buff")
int snmfnat_blk_sparse(snmfnat_ctx*, const double*, const double*, const double*, int, int, const snmfnat_params*, double*, double*) SN_STUB("snmfnat_blk_sparse")
int snmfnat_stream_create(snmfnat_ctx*, const snmfnat_params*, const double*, const double*, const double*, const double*, int, const double*, const double*, int, const double*, const double*, snmfnat_stream**) SN_STUB("snmfnat_stream_create")
int snmfnat_stream_destroy(snmfnat_stream*) SN_STUB("snmfnat_stream_destroy")
int snmfnat_stream_step(snmfnat_stream*, const double*, int, const double*, double*, double*, double*) SN_STUB("snmfnat_stream_step")
int snmfnat_stream_get(snmfnat_stream*, const char*, double*, int64_t) SN_STUB("snmfnat_stream_get")
int snmfnat_stream_set(snmfnat_stream*, const char*, const double*, int64_t) SN_STUB("snmfnat_stream_set")
}
