/**
 * \file test_ipc_benchmark.cc
 * \brief Co-located worker + server benchmark (one joint process per GPU / host).
 *
 * The reference's test_ipc_benchmark (tests/test_ipc_benchmark.cc:24-266) measures its
 * same-host shortcut: worker and server exchange values through a named POSIX shm
 * segment so the push carries only the meta. Here every process runs DMLC_ROLE=joint:
 * values are allocated from the van's peer-mappable memory (HBM on the nvl van, a shm
 * arena on the shm van), pushes are one-sided writes, and the push to the *co-located*
 * server (same process) needs no IPC mapping at all.
 *   usage: test_ipc_benchmark [len=1024000] [rounds=100]     env: NUM_KEY_PER_SERVER (10)
 * Launch with JOINT=1 scripts/local.sh N N build/test_ipc_benchmark (N joint processes).
 */
#include <chrono>
#include "ps/ps.h"
#if PS_USE_CUDA
#include <cuda_runtime.h>
#endif
using namespace ps;

int main(int argc, char* argv[]) {
  const int len = argc > 1 ? atoi(argv[1]) : 1024000;
  const int rounds = argc > 2 ? atoi(argv[2]) : 100;
  const int kps = GetEnv("NUM_KEY_PER_SERVER", 10);
  const std::string role_str = GetEnvStr("DMLC_ROLE", "joint");
  const Node::Role role = GetRole(role_str);
  if (role == Node::SCHEDULER) {
    StartPS(0, role, -1, true);
    Finalize(0, role, true);
    return 0;
  }
  CHECK(role == Node::JOINT) << "test_ipc_benchmark runs DMLC_ROLE=joint (and a scheduler)";
#if PS_USE_CUDA
  if (GetEnvStr("PS_VAN_TYPE") == "nvl") cudaSetDevice(GetEnv("PS_CUDA_DEVICE", 0));
#endif
  StartPS(0, role, GetEnv("DMLC_RANK", -1), true);
  Van* wvan = Postoffice::GetWorker()->van();
  const bool gpu = wvan->GetType() == "nvl";

  // server half: the landing buffer of the first push is the store
  std::mutex mu;
  std::unordered_map<Key, KVPairs<char>> store;
  KVServer<char> server(0);
  server.set_request_handle([&](const KVMeta& req, const KVPairs<char>& d, KVServer<char>* s) {
    const Key key = d.keys[0];
    if (req.push) {
      {
        std::lock_guard<std::mutex> lk(mu);
        if (!store.count(key)) {
          auto& slot = store[key];
          slot.keys.CopyFrom(d.keys);
          slot.lens.CopyFrom(d.lens);
          slot.vals = d.vals;
        }
      }
      s->Response(req);
    } else {
      KVPairs<char> res;
      {
        std::lock_guard<std::mutex> lk(mu);
        res = store.at(key);
      }
      s->Response(req, res);
    }
  });

  // worker half
  // neither the handler above nor this worker's callbacks wait for the network
  const bool inline_dispatch = GetEnv("BENCHMARK_INLINE", 1) != 0;
  server.set_inline_dispatch(inline_dispatch);
  KVWorker<char> kv(0, 0);
  kv.set_inline_dispatch(inline_dispatch);
  const auto& ranges = Postoffice::GetWorker()->GetServerKeyRanges();
  const int S = static_cast<int>(ranges.size());
  const int total = S * kps;
  std::vector<SArray<Key>> keys(total);
  std::vector<SArray<char>> vals(total);
  std::vector<SArray<int>> lens(total);
  const int dev = gpu ? wvan->my_node().dev_id : 0;
  for (int k = 0; k < total; ++k) {
    keys[k] = SArray<Key>(1, static_cast<Key>(ranges[k % S].begin() + k));
    lens[k] = SArray<int>(1, len);
    char* p = static_cast<char*>(wvan->AllocExportable(len));
    CHECK(p);
    vals[k].reset(p, len, [](char*) {}, gpu ? GPU : CPU, dev, gpu ? GPU : CPU, dev);
  }
  for (int k = 0; k < total; ++k) kv.Wait(kv.ZPush(keys[k], vals[k], lens[k]));
  Postoffice::GetWorker()->Barrier(0, kWorkerGroup);
  std::vector<int> ts;
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < rounds; ++r) {
    for (int k = 0; k < total; ++k) {
      ts.push_back(kv.ZPush(keys[k], vals[k], lens[k]));
      ts.push_back(kv.ZPull(keys[k], &vals[k], &lens[k]));
    }
    for (int t : ts) kv.Wait(t);
    ts.clear();
  }
  const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
  LL << "[joint " << Postoffice::GetWorker()->my_rank() << "]\tApplication goodput: "
     << 8.0 * len * total * rounds / ns << " Gbps.\tAvg latency = " << ns / rounds / total / 1000.0
     << " us per key (" << wvan->GetType() << " van)";
  Finalize(0, role, true);
  return 0;
}
