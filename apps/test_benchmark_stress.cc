/**
 * \file test_benchmark_stress.cc
 * \brief Traffic-pattern stress test: N threads share ONE KVWorker in a joint process.
 *
 * Emulates the four session-level collectives the reference stresses
 * (tests/test_benchmark_stress.cc:249-263): per minibatch every session (thread) issues
 *   DataScatter : ZPush to every *remote* node          (key family 0)
 *   Gather      : ZPull from every remote node          (key family 1, shared with Scatter)
 *   Scatter     : ZPush to every remote node            (key family 1)
 *   DenseReduce : ZPush + ZPull against every node      (key family 2)
 * and waits for all of them. Each thread verifies a stamp in the first 8 bytes of what
 * it pulls (DEBUG_MODE=1), so lost or crossed messages are detected, and prints per-phase
 * times. Roles: scheduler, or joint (BYTEPS_NODE_ID = this node's index, defaults to
 * DMLC_RANK).   usage: test_benchmark_stress [len=30720000] [repeat=20]   env: BENCHMARK_NTHREAD (8)
 *
 * On the nvl van (or STRESS_GPU=1) every buffer lives in HBM (the reference's GPU mode,
 * tests/test_benchmark_stress.cc:249-432): a minibatch's payload is generated on the device
 * (k_fill_u32, seeded by session / minibatch / destination), travels over NVLink, and what the
 * Gather and DenseReduce phases pull back is CHECKSUMMED in full on the device (k_checksum_u32)
 * against the checksum of what was pushed — not just an 8-byte stamp. Works under fault
 * injection too: PS_DROP_MSG=5 PS_RESEND=1.
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
struct Buf {
  SArray<Key> key;
  SArray<char> val;
  SArray<int> len;
};

// the key of (family, src session, dst node) lives on server `dst`
Key MakeKey(const std::vector<Range>& ranges, int family, int session, int dst, int sessions) {
  return static_cast<Key>(ranges[dst].begin() + 1 + family * 1000000 + dst * sessions + session);
}
}  // namespace

int main(int argc, char* argv[]) {
  const int len = argc > 1 ? atoi(argv[1]) : 1024000 * 30;
  const int repeat = argc > 2 ? atoi(argv[2]) : 20;
  const int nthread = GetEnv("BENCHMARK_NTHREAD", 8);
  const bool debug = Environment::Get()->find("DEBUG_MODE") != nullptr;
  // with fault injection a pull can overtake the (retransmitted) push it follows: DenseReduce, which
  // issues both without waiting in between, is then only exercised, not checked byte for byte
  const bool lossy = GetEnv("PS_DROP_MSG", 0) > 0;
  const std::string role_str = CHECK_NOTNULL(Environment::Get()->find("DMLC_ROLE"));
  const Node::Role role = GetRole(role_str);
  if (role == Node::SCHEDULER) {
    StartPS(0, role, -1, true);
    Finalize(0, role, true);
    return 0;
  }
  CHECK(role == Node::JOINT) << "the stress test runs joint nodes";
  const int node_id = GetEnv("BYTEPS_NODE_ID", GetEnv("DMLC_RANK", 0));
#if PS_USE_CUDA
  if (GetEnvStr("PS_VAN_TYPE") == "nvl") cudaSetDevice(GetEnv("PS_CUDA_DEVICE", 0));
#endif
  StartPS(0, role, node_id, true);
  Van* van = Postoffice::GetWorker()->van();
  bool gpu = van->GetType() == "nvl";
#if !PS_USE_CUDA
  gpu = false;
#endif
  const int gpu_dev = gpu ? van->my_node().dev_id : 0;

  std::mutex mu;
  std::unordered_map<Key, SArray<char>> store;
  KVServer<char> server(0);
  server.set_request_handle([&](const KVMeta& req, const KVPairs<char>& d, KVServer<char>* s) {
    const Key key = d.keys[0];
    if (req.push && gpu) {
      // the landing slot in HBM is the store: a pull is answered from it, byte for byte
      {
        std::lock_guard<std::mutex> lk(mu);
        store[key] = d.vals;
      }
      s->Response(req);
      return;
    }
    if (req.push) {
      {
        std::lock_guard<std::mutex> lk(mu);
        SArray<char>& slot = store[key];
        if (slot.size() != d.vals.size()) slot = SArray<char>(d.vals.size(), 0);
        // accumulate floats like a reducing server would (DenseReduce), cheaply: first 64 B
        const size_t n = std::min<size_t>(16, d.vals.size() / 4);
        float* dst = reinterpret_cast<float*>(slot.data());
        const float* src = reinterpret_cast<const float*>(d.vals.data());
        for (size_t i = 2; i < n; ++i) dst[i] += src[i];
        memcpy(slot.data(), d.vals.data(), std::min<size_t>(8, d.vals.size()));  // stamp
      }
      s->Response(req);
    } else {
      KVPairs<char> res;
      res.keys = d.keys;
      {
        std::lock_guard<std::mutex> lk(mu);
        auto it = store.find(key);
        if (it == store.end()) {
          // only possible when messages are being dropped: the push this pull follows was lost and its
          // retransmission has not arrived yet (a resend does not keep the order of the stream)
          CHECK(lossy) << "pull before push of key " << key;
          SArray<char>& slot = store[key];
          slot = gpu ? SArray<char>() : SArray<char>(static_cast<size_t>(req.val_len > 0 ? req.val_len : len), 0);
          it = store.find(key);
        }
        if (it->second.size() == 0) {  // (device mode, lossy: nothing has landed yet — answer without values)
          s->Response(req);
          return;
        }
        res.vals = it->second;
      }
      res.lens = SArray<int>(1, static_cast<int>(res.vals.size()));
      s->Response(req, res);
    }
  });

  // neither the handler above nor this worker's callbacks wait for the network
  const bool inline_dispatch = GetEnv("BENCHMARK_INLINE", 1) != 0;
  server.set_inline_dispatch(inline_dispatch);
  KVWorker<char> kv(0, 0);
  kv.set_inline_dispatch(inline_dispatch);
  const auto& ranges = Postoffice::GetWorker()->GetServerKeyRanges();
  const int nodes = static_cast<int>(ranges.size());
  const int sessions = nthread * nodes;
  std::atomic<int> failures{0};
  std::vector<std::thread> threads;
  for (int tid = 0; tid < nthread; ++tid) {
    threads.emplace_back([&, tid] {
      const int session = node_id * nthread + tid;
      // one buffer per (family, dst)
#if PS_USE_CUDA
      cudaStream_t stream = nullptr;
      unsigned long long* sum_dev = nullptr;
      if (gpu) {
        CHECK(cudaSetDevice(gpu_dev) == cudaSuccess);
        CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess);
        CHECK(cudaMalloc(reinterpret_cast<void**>(&sum_dev), 16) == cudaSuccess);
      }
      // fill a device buffer with this minibatch's pattern / checksum all of it
      auto gpu_fill = [&](Buf& b, uint32_t seed) {
        const ps_stream_t st = reinterpret_cast<ps_stream_t>(stream);
        CHECK_EQ(ps_launch_fill_u32(b.val.data(), static_cast<size_t>(len) / 4, seed, st), 0);
        CHECK(cudaStreamSynchronize(stream) == cudaSuccess);
      };
      auto gpu_sum = [&](const Buf& b) {
        unsigned long long h = 0;
        const ps_stream_t st = reinterpret_cast<ps_stream_t>(stream);
        CHECK_EQ(ps_launch_checksum_u32(b.val.data(), static_cast<size_t>(len) / 4, sum_dev, st), 0);
        CHECK(cudaMemcpyAsync(&h, sum_dev, 8, cudaMemcpyDeviceToHost, stream) == cudaSuccess);
        CHECK(cudaStreamSynchronize(stream) == cudaSuccess);
        return h;
      };
      auto gpu_clear = [&](Buf& b) {
        CHECK(cudaMemsetAsync(b.val.data(), 0, static_cast<size_t>(len), stream) == cudaSuccess);
        CHECK(cudaStreamSynchronize(stream) == cudaSuccess);
      };
#endif
      auto make = [&](int family, int dst) {
        Buf b;
        b.key = SArray<Key>(1, MakeKey(ranges, family, session, dst, sessions));
        if (gpu) {
          char* p = static_cast<char*>(van->AllocExportable(static_cast<size_t>(len)));
          CHECK(p) << "out of device memory";
          b.val.reset(p, static_cast<size_t>(len), [](char*) {}, GPU, gpu_dev, GPU, gpu_dev);
        } else {
          b.val = SArray<char>(static_cast<size_t>(len), 1);
        }
        b.len = SArray<int>(1, len);
        return b;
      };
      std::vector<Buf> data_scatter, gs, dense;
      for (int d = 0; d < nodes; ++d) {
        data_scatter.push_back(make(0, d));
        gs.push_back(make(1, d));
        dense.push_back(make(2, d));
      }
      std::vector<unsigned long long> want_gs(static_cast<size_t>(nodes), 0), want_dense(static_cast<size_t>(nodes), 0);
      auto stamp = [&](Buf& b, int mb, int dst = 0, unsigned long long* want = nullptr) {
#if PS_USE_CUDA
        if (gpu) {
          gpu_fill(b, static_cast<uint32_t>(session * 1000003 + mb * 131 + dst));
          if (want) *want = gpu_sum(b);
          return;
        }
#endif
        (void)dst;
        (void)want;
        uint64_t s = (static_cast<uint64_t>(session) << 32) | static_cast<uint32_t>(mb);
        memcpy(b.val.data(), &s, 8);
      };
      double t_ds = 0, t_g = 0, t_s = 0, t_d = 0;
      auto now = [] { return std::chrono::steady_clock::now(); };
      auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
      std::vector<int> ts;
      auto drain = [&] {
        for (int t : ts) kv.Wait(t);
        ts.clear();
      };
      for (int mb = 0; mb < repeat; ++mb) {
        auto a = now();
        for (int d = 0; d < nodes; ++d) {
          if (d == node_id && nodes > 1) continue;
          stamp(data_scatter[d], mb);
          ts.push_back(kv.ZPush(data_scatter[d].key, data_scatter[d].val, data_scatter[d].len));
        }
        drain();
        auto b = now();
        t_ds += ms(a, b);
        for (int d = 0; d < nodes; ++d) {  // Scatter (push) ...
          if (d == node_id && nodes > 1) continue;
          stamp(gs[d], mb, d, &want_gs[static_cast<size_t>(d)]);
          ts.push_back(kv.ZPush(gs[d].key, gs[d].val, gs[d].len));
        }
        drain();
        auto c = now();
        t_s += ms(b, c);
        for (int d = 0; d < nodes; ++d) {  // ... then Gather (pull) the same keys back
          if (d == node_id && nodes > 1) continue;
#if PS_USE_CUDA
          if (gpu) gpu_clear(gs[d]);
#endif
          if (!gpu) memset(gs[d].val.data(), 0, 8);
          ts.push_back(kv.ZPull(gs[d].key, &gs[d].val, &gs[d].len));
        }
        drain();
        auto e = now();
        t_g += ms(c, e);
#if PS_USE_CUDA
        if (gpu) {  // every byte that came back over NVLink is checked, every minibatch
          for (int d = 0; d < nodes; ++d) {
            if (d == node_id && nodes > 1) continue;
            if (gpu_sum(gs[d]) != want_gs[static_cast<size_t>(d)]) {
              ++failures;
              LOG(ERROR) << "session " << session << " minibatch " << mb << ": gather from node " << d << " is corrupt";
            }
          }
        }
#endif
        if (debug && !gpu) {
          for (int d = 0; d < nodes; ++d) {
            if (d == node_id && nodes > 1) continue;
            uint64_t s;
            memcpy(&s, gs[d].val.data(), 8);
            if (s != ((static_cast<uint64_t>(session) << 32) | static_cast<uint32_t>(mb))) ++failures;
          }
        }
        for (int d = 0; d < nodes; ++d) {  // DenseReduce: push + pull against every node
          stamp(dense[d], mb, d + 100, &want_dense[static_cast<size_t>(d)]);
          ts.push_back(kv.ZPush(dense[d].key, dense[d].val, dense[d].len));
          ts.push_back(kv.ZPull(dense[d].key, &dense[d].val, &dense[d].len));
        }
        drain();
        t_d += ms(e, now());
#if PS_USE_CUDA
        if (gpu && !lossy) {
          for (int d = 0; d < nodes; ++d) {
            if (gpu_sum(dense[d]) != want_dense[static_cast<size_t>(d)]) {
              ++failures;
              LOG(ERROR) << "session " << session << " minibatch " << mb << ": dense_reduce with node " << d
                         << " is corrupt";
            }
          }
        }
#endif
      }
      LL << "[node " << node_id << " session " << tid << "] per minibatch: data_scatter "
         << t_ds / repeat << " ms, scatter " << t_s / repeat << " ms, gather " << t_g / repeat
         << " ms, dense_reduce " << t_d / repeat << " ms";
    });
  }
  for (auto& t : threads) t.join();
  LL << (failures.load() ? "test_benchmark_stress FAILED" : "test_benchmark_stress PASSED")
     << " on node " << node_id << (gpu ? " (HBM buffers, full-payload checksums)" : "");
  Finalize(0, role, true);
  return failures.load() ? 1 : 0;
}
