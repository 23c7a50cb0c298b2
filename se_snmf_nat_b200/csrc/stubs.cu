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
}
