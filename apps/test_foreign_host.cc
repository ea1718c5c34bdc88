/**
 * \file test_foreign_host.cc
 * \brief A one-sided van with peers on ANOTHER host: nothing can be mapped, so values travel in
 *        socket frames — device memory staged through the host on both ends — and must still
 *        arrive bit for bit. Worker: push a pattern from exportable memory, pull it back into
 *        exportable memory (plain and fused push-pull), compare. Server: store[key] += vals.
 *
 * Run the worker with a DMLC_NODE_HOST different from the server's (127.0.0.2 against 127.0.0.1
 * works on one Linux box). PS_TEST_STAGE_ARENA=1 makes the shm van treat its arena like device
 * memory, which exercises the staging copies without a GPU; with PS_VAN_TYPE=nvl the values
 * live in HBM (TEST_FOREIGN_GPU=1).
 */
#include <cmath>
#include <unordered_map>
#include "ps/ps.h"
#if PS_USE_CUDA
#include <cuda_runtime.h>
#endif
using namespace ps;

int main() {
  const std::string role_str = CHECK_NOTNULL(Environment::Get()->find("DMLC_ROLE"));
  const Node::Role role = GetRole(role_str);
  const bool gpu = GetEnv("TEST_FOREIGN_GPU", 0) != 0;
#if PS_USE_CUDA
  if (gpu) cudaSetDevice(GetEnv("PS_CUDA_DEVICE", 0));
#endif
  StartPS(0, role, -1, true);
  if (role == Node::SCHEDULER) {
    Finalize(0, role, true);
    return 0;
  }
  if (role == Node::SERVER) {
    // store[key] += vals, one vector per key (KVServerDefaultHandle keeps one value per key)
    std::unordered_map<Key, std::vector<float>> store;
    KVServer<float> server(0);
    server.set_request_handle([&store](const KVMeta& req, const KVPairs<float>& d, KVServer<float>* s) {
      CHECK_EQ(d.keys.size(), (size_t)1);
      std::vector<float>& v = store[d.keys[0]];
      if (req.push) {
        CHECK(!d.vals.on_gpu()) << "values from another host arrive in host memory";
        if (v.size() < d.vals.size()) v.resize(d.vals.size(), 0.f);
        for (size_t i = 0; i < d.vals.size(); ++i) v[i] += d.vals[i];
      }
      if (req.push && !req.pull) {
        s->Response(req);
        return;
      }
      KVPairs<float> res;
      res.keys = d.keys;
      res.vals.CopyFrom(v.data(), v.size());
      res.lens = SArray<int>(1, static_cast<int>(v.size()));
      s->Response(req, res);
    });
    Finalize(0, role, true);
    return 0;
  }
  Van* van = Postoffice::Get()->van();
  KVWorker<float> kv(0, 0);
  const int n = GetEnv("TEST_FOREIGN_LEN", 100000);
  std::vector<float> pattern(static_cast<size_t>(n)), got(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) pattern[static_cast<size_t>(i)] = static_cast<float>((i * 31 + 7) % 1000) * 0.25f;
  auto buffer = [&]() {
    float* p = static_cast<float*>(van->AllocExportable(sizeof(float) * static_cast<size_t>(n)));
    CHECK(p) << "this test needs a van with exportable memory (shm / nvl)";
    SArray<float> a;
    a.reset(p, static_cast<size_t>(n), [](float*) {}, gpu ? GPU : CPU, gpu ? van->my_node().dev_id : 0,
            gpu ? GPU : CPU, gpu ? van->my_node().dev_id : 0);
    return a;
  };
  auto upload = [&](SArray<float>& dst, const std::vector<float>& src) {
#if PS_USE_CUDA
    if (gpu) {
      CHECK(cudaMemcpy(dst.data(), src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess);
      return;
    }
#endif
    memcpy(dst.data(), src.data(), src.size() * sizeof(float));
  };
  auto download = [&](std::vector<float>& dst, const SArray<float>& src) {
#if PS_USE_CUDA
    if (gpu) {
      CHECK(cudaMemcpy(dst.data(), src.data(), dst.size() * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess);
      return;
    }
#endif
    memcpy(dst.data(), src.data(), dst.size() * sizeof(float));
  };
  SArray<Key> keys(std::vector<Key>{3});
  SArray<float> vals = buffer(), out = buffer(), out2 = buffer();
  upload(vals, pattern);
  std::vector<float> zeros(static_cast<size_t>(n), -1.f);
  upload(out, zeros);
  upload(out2, zeros);
  kv.Wait(kv.ZPush(keys, vals));
  kv.Wait(kv.ZPull(keys, &out));
  download(got, out);
  for (int i = 0; i < n; ++i) {
    CHECK_EQ(got[static_cast<size_t>(i)], pattern[static_cast<size_t>(i)]) << "pull, element " << i;
  }
  // fused push-pull: the store becomes 2 x pattern and comes back in the same reply
  kv.Wait(kv.ZPushPull(keys, vals, &out2));
  download(got, out2);
  for (int i = 0; i < n; ++i) {
    CHECK_EQ(got[static_cast<size_t>(i)], 2.f * pattern[static_cast<size_t>(i)]) << "push-pull, element " << i;
  }
  std::vector<std::pair<std::string, uint64_t>> stats;
  van->TransportStats(&stats);
  uint64_t onesided = 0, staged = 0;
  for (auto& kvp : stats) {
    if (kvp.first == "onesided_copies") onesided = kvp.second;
    if (kvp.first == "staged_copies") staged = kvp.second;
  }
  LL << "test_foreign_host PASSED: one-sided copies " << onesided << ", staged copies " << staged;
  Finalize(0, role, true);
  return 0;
}
