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
 * Launch with JOINT=1 scripts/local.sh N N build/test_ipc_benchmark (N joint processes);
 * extra DMLC_ROLE=server processes are plain (non-co-located) servers.
 *
 *   BYTEPS_ENABLE_MIXED_MODE=1  keys are spread over co-located and non-co-located servers with
 *                               the reference's load formula (tests/test_ipc_benchmark.cc:144-166)
 *   IPC_NVLS_PULL=1             the values live in ONE symmetric buffer (Van::AllocSymmetric) and a
 *                               pull is answered THROUGH THE SWITCH: the server stores the value once
 *                               to the multicast address (multimem.st) and it lands in every worker's
 *                               copy — server egress 1x the payload for W workers (BASELINE.json
 *                               config 5). Needs the nvl van on an NVSwitch box.
 *   IPC_VERIFY=1                every worker checks the bytes it pulled in the last round
 */
#include <chrono>
#include <cstring>
#include "ps/ps.h"
#if PS_USE_CUDA
#include <cuda_runtime.h>
#include "kernels/ps_kernels.h"
#endif
using namespace ps;

namespace {

/*! \brief server that key `seed` of `total` lives on (reference tests/test_ipc_benchmark.cc:144-166) */
int AllocateServer(int seed, int total, int num_servers, int num_workers, bool mixed) {
  if (!mixed) return seed % num_servers;
  const int noncoloc = num_servers - num_workers, coloc = num_workers;
  if (noncoloc <= 0) return seed % num_servers;
  const double ratio = (2.0 * noncoloc * (num_workers - 1)) /
                       (static_cast<double>(num_workers) * (num_workers + noncoloc) - 2.0 * noncoloc);
  if (seed < ratio * total) return seed % noncoloc;
  return noncoloc + (seed % coloc);
}

}  // namespace

int main(int argc, char* argv[]) {
  const int len = argc > 1 ? atoi(argv[1]) : 1024000;
  const int rounds = argc > 2 ? atoi(argv[2]) : 100;
  const int kps = GetEnv("NUM_KEY_PER_SERVER", 10);
  const bool mixed = GetEnv("BYTEPS_ENABLE_MIXED_MODE", 0) != 0;
  const bool nvls = GetEnv("IPC_NVLS_PULL", 0) != 0;
  const bool verify = GetEnv("IPC_VERIFY", 0) != 0;
  const std::string role_str = GetEnvStr("DMLC_ROLE", "joint");
  const Node::Role role = GetRole(role_str);
  if (role == Node::SCHEDULER) {
    StartPS(0, role, -1, true);
    Finalize(0, role, true);
    return 0;
  }
  CHECK(role == Node::JOINT || role == Node::SERVER)
      << "test_ipc_benchmark runs DMLC_ROLE=joint processes (plus plain servers and a scheduler)";
#if PS_USE_CUDA
  if (GetEnvStr("PS_VAN_TYPE") == "nvl") cudaSetDevice(GetEnv("PS_CUDA_DEVICE", 0));
#endif
  StartPS(0, role, GetEnv("DMLC_RANK", -1), true);
  Van* svan = Postoffice::GetServer()->van();
  const bool gpu = svan->GetType() == "nvl";
  const int W = NumWorkers(), S = NumServers();
  const int total = S * kps;
  const size_t stride = (static_cast<size_t>(len) + 255) & ~size_t(255);  // slot of a key in the symmetric buffer

  // NVLS pull: one symmetric buffer holds every key's value in every process
  SymmetricBuffer symm;
  if (nvls) {
    CHECK(svan->AllocSymmetric("ipc_vals", stride * static_cast<size_t>(total), &symm))
        << "IPC_NVLS_PULL needs a van with symmetric memory (nvl / shm)";
    if (!symm.mc) LOG(WARNING) << "no NVSwitch multicast here: pulls are answered peer by peer";
  }
  // (multicast is all-or-nothing across the members, so every process takes the same branch)
  const bool mc_pull = nvls && symm.mc != nullptr;
  const bool vals_in_symm = mc_pull || (nvls && !gpu);

  // server half: the landing buffer of the first push is the store
  std::mutex mu;
  std::unordered_map<Key, KVPairs<char>> store;
  std::unordered_map<Key, int> pulls_served;
  std::atomic<uint64_t> mc_fanouts{0};
  KVServer<char> server(0);
  const auto& ranges = Postoffice::GetServer()->GetServerKeyRanges();
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
      return;
    }
    KVPairs<char> res;
    int nth = 0;
    {
      std::lock_guard<std::mutex> lk(mu);
      res = store.at(key);
      nth = pulls_served[key]++;
    }
    (void)nth;  // (only the multicast path of a CUDA build looks at it)
#if PS_USE_CUDA
    if (mc_pull && gpu && req.mem.region == kSymmetricRegion) {
      // the first pull of a round publishes the value to ALL workers with one multimem.st stream;
      // this reply and those to the other workers only have to wait for that kernel
      if (nth % W == 0) {
        const int rc = ps_launch_copy_multicast(static_cast<char*>(symm.mc) + req.mem.offset, res.vals.data(),
                                                (res.vals.size() + 15) & ~size_t(15), 0, nullptr,
                                                reinterpret_cast<ps_stream_t>(svan->DataStream()));
        CHECK_EQ(rc, 0) << "multicast copy failed: " << cudaGetErrorString(static_cast<cudaError_t>(rc));
        ++mc_fanouts;
      }
      SendOpts placed;
      placed.codec = kCodecPlaced;
      s->Response(req, res, placed);
      return;
    }
#endif
    s->Response(req, res);
  });
  const bool inline_dispatch = GetEnv("BENCHMARK_INLINE", 1) != 0;
  server.set_inline_dispatch(inline_dispatch);

  if (role == Node::SERVER) {  // a plain server only serves
    Finalize(0, role, true);
    return 0;
  }

  // worker half
  Van* wvan = Postoffice::GetWorker()->van();
  KVWorker<char> kv(0, 0);
  kv.set_inline_dispatch(inline_dispatch);
  std::vector<SArray<Key>> keys(total);
  std::vector<SArray<char>> vals(total);
  std::vector<SArray<int>> lens(total);
  const int dev = gpu ? wvan->my_node().dev_id : 0;
  for (int k = 0; k < total; ++k) {
    const int srv = AllocateServer(k, total, S, W, mixed);
    keys[k] = SArray<Key>(1, static_cast<Key>(ranges[srv].begin() + k));
    lens[k] = SArray<int>(1, len);
    char* p = vals_in_symm ? static_cast<char*>(symm.local) + stride * static_cast<size_t>(k)
                           : static_cast<char*>(wvan->AllocExportable(len));
    CHECK(p);
    vals[k].reset(p, len, [](char*) {}, gpu ? GPU : CPU, dev, gpu ? GPU : CPU, dev);
  }
  // every worker pushes the same bytes for a key (a pull is answered from the first pusher's slot)
  std::vector<char> pattern(static_cast<size_t>(len));
  auto fill = [&](int k) {
    for (int i = 0; i < len; ++i) pattern[static_cast<size_t>(i)] = static_cast<char>((i * 7 + k * 13 + 5) & 0xff);
#if PS_USE_CUDA
    if (gpu) {
      CHECK(cudaMemcpy(vals[k].data(), pattern.data(), pattern.size(), cudaMemcpyHostToDevice) == cudaSuccess);
      return;
    }
#endif
    memcpy(vals[k].data(), pattern.data(), pattern.size());
  };
  for (int k = 0; k < total; ++k) fill(k);
  auto pull = [&](int k) {
    if (!mc_pull) return kv.ZPull(keys[k], &vals[k], &lens[k]);
    SendOpts o;  // the server knows the destination as an offset inside the symmetric buffer
    o.dest_mem.region = kSymmetricRegion;
    o.dest_mem.offset = stride * static_cast<uint64_t>(k);
    o.dest_mem.bytes = static_cast<uint64_t>(len);
    return kv.ZPull(keys[k], &vals[k], &lens[k], 0, nullptr, o);
  };
  for (int k = 0; k < total; ++k) kv.Wait(kv.ZPush(keys[k], vals[k], lens[k]));
  Postoffice::GetWorker()->Barrier(0, kWorkerGroup);
  std::vector<int> ts;
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < rounds; ++r) {
    for (int k = 0; k < total; ++k) {
      ts.push_back(kv.ZPush(keys[k], vals[k], lens[k]));
      ts.push_back(pull(k));
    }
    for (int t : ts) kv.Wait(t);
    ts.clear();
    // with the multicast fan-out a reply overwrites EVERY worker's copy: nobody may push the next
    // round's value of a key while a slower worker has not read this round's yet
    if (mc_pull) Postoffice::GetWorker()->Barrier(0, kWorkerGroup);
  }
  const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
  bool ok = true;
  if (verify) {
    std::vector<char> got(static_cast<size_t>(len));
    for (int k = 0; k < total && ok; ++k) {
      for (int i = 0; i < len; ++i) pattern[static_cast<size_t>(i)] = static_cast<char>((i * 7 + k * 13 + 5) & 0xff);
#if PS_USE_CUDA
      if (gpu) {
        CHECK(cudaMemcpy(got.data(), vals[k].data(), got.size(), cudaMemcpyDeviceToHost) == cudaSuccess);
      } else
#endif
      {
        memcpy(got.data(), vals[k].data(), got.size());
      }
      ok = memcmp(got.data(), pattern.data(), got.size()) == 0;
      if (!ok) LOG(ERROR) << "key " << k << " came back different from what was pushed";
    }
  }
  LL << "[joint " << Postoffice::GetWorker()->my_rank() << "]\tApplication goodput: "
     << 8.0 * len * total * rounds / ns << " Gbps.\tAvg latency = " << ns / rounds / total / 1000.0
     << " us per key (" << wvan->GetType() << " van"
     << (nvls ? (mc_pull ? ", NVLS multicast pull" : ", symmetric buffer, unicast pull") : "")
     << (mixed ? ", mixed mode" : "") << ", server multicast fan-outs " << mc_fanouts.load() << ")"
     << (verify ? (ok ? " VERIFIED" : " VERIFY FAILED") : "");
  Finalize(0, role, true);
  return ok ? 0 : 1;
}
