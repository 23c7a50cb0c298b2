// Corpus enhancement on several GPUs of one node from ONE host process: the utterances (or chains of utterances) of a
// corpus are independent (filewise_run_IS16.m:54-186; Do_MultiBatch_IS16_20160324_CHiME4.m:202-208 walks them one by
// one), so they are split over the devices longest-processing-time first and every device runs snmfnat_enhance_batch on
// its share from its own host thread.  No data-path collective.  (One process per GPU with torch.distributed, as bench.py
// does, is the other way to drive several devices; this entry is what a single MATLAB / Octave process binds.)
#include <algorithm>
#include <mutex>
#include <numeric>
#include <thread>
#include "common.cuh"

using namespace snmfnat;

extern "C" int snmfnat_enhance_batch_multi(const int* devices, int n_dev, const snmfnat_params* p, const double* win_stft,
                                           const double* win_istft, const double* B_x, const double* B_d, int n2, int n_utt,
                                           const int16_t* const* pcm, const int64_t* len, const int32_t* chain_id,
                                           const double* h_init, const double* Ad_blk_init, int64_t ad_stride,
                                           int16_t* const* out) {
  SN_API_BEGIN
  SN_REQUIRE(devices && n_dev >= 1 && p && pcm && len && out && n_utt > 0, SNMFNAT_EINVAL, "bad argument");
  // units: single utterances, or chains (equal chain_id >= 0) that must stay together and in order
  std::vector<std::vector<int>> units;
  {
    std::vector<std::pair<int, int>> seen;
    for (int u = 0; u < n_utt; ++u) {
      const int cid = chain_id ? chain_id[u] : -1;
      int at = -1;
      if (cid >= 0)
        for (auto& pr : seen)
          if (pr.first == cid) at = pr.second;
      if (at < 0) {
        at = (int)units.size();
        units.emplace_back();
        if (cid >= 0) seen.emplace_back(cid, at);
      }
      units[at].push_back(u);
    }
  }
  const int shift = p->frameshift > 0 ? p->frameshift : 1;
  std::vector<long long> cost(units.size(), 0);
  for (size_t j = 0; j < units.size(); ++j)
    for (int u : units[j]) cost[j] += len[u] / shift + p->delay + 1;          // hops of the file, filewise_run_IS16.m:102-123
  std::vector<int> order(units.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  std::vector<std::vector<int>> share(n_dev);          // utterance indices per device
  std::vector<long long> load(n_dev, 0);
  std::vector<std::vector<int>> unit_of_dev(n_dev);
  for (int j : order) {
    const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    load[d] += cost[j];
    unit_of_dev[d].push_back(j);
  }
  for (int d = 0; d < n_dev; ++d) {
    std::sort(unit_of_dev[d].begin(), unit_of_dev[d].end());                    // corpus order inside a device
    for (int j : unit_of_dev[d])
      for (int u : units[j]) share[d].push_back(u);
  }
  std::mutex mu;
  int first_rc = SNMFNAT_OK;
  std::string first_msg;
  auto work = [&](int d) {
    const std::vector<int>& idx = share[d];
    if (idx.empty()) return;
    snmfnat_ctx* ctx = nullptr;
    int rc = snmfnat_ctx_create(devices[d], &ctx);
    if (rc == SNMFNAT_OK) {
      const size_t n = idx.size();
      std::vector<const int16_t*> in(n);
      std::vector<int16_t*> o(n);
      std::vector<int64_t> ln(n);
      std::vector<int32_t> ch(n, -1);
      std::vector<double> ad;
      if (Ad_blk_init) ad.resize(n * (size_t)ad_stride);
      for (size_t i = 0; i < n; ++i) {
        in[i] = pcm[idx[i]];
        o[i] = out[idx[i]];
        ln[i] = len[idx[i]];
        if (chain_id) ch[i] = chain_id[idx[i]];
        if (Ad_blk_init) std::copy(Ad_blk_init + (size_t)idx[i] * ad_stride, Ad_blk_init + (size_t)(idx[i] + 1) * ad_stride, ad.begin() + i * (size_t)ad_stride);
      }
      rc = snmfnat_enhance_batch(ctx, p, win_stft, win_istft, B_x, B_d, n2, (int)n, in.data(), ln.data(),
                                 chain_id ? ch.data() : nullptr, h_init, Ad_blk_init ? ad.data() : nullptr, ad_stride, o.data());
    }
    if (rc != SNMFNAT_OK) {
      std::lock_guard<std::mutex> lk(mu);
      if (first_rc == SNMFNAT_OK) {
        first_rc = rc;
        first_msg = "device " + std::to_string(devices[d]) + ": " + snmfnat_last_error(ctx);   // thread-local message of this thread
      }
    }
    if (ctx) snmfnat_ctx_destroy(ctx);
  };
  std::vector<std::thread> th;
  for (int d = 1; d < n_dev; ++d) th.emplace_back(work, d);
  work(0);
  for (auto& t : th) t.join();
  if (first_rc != SNMFNAT_OK) fail(first_rc, "%s", first_msg.c_str());
  SN_API_END
}
