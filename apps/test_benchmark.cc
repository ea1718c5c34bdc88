/**
 * \file test_benchmark.cc
 * \brief Push/pull throughput + latency benchmark (the framework's headline test).
 *
 * Same command line and environment contract as the reference program
 * (tests/test_benchmark.cc:25-30 modes, :131-203 server handler, :322-397 timing
 * loop, :489-553 main):
 *   argv:  [1] bytes per value (1024000)  [2] repeat (10, mode 0 only)
 *          [3] mode 0=PUSH_THEN_PULL 1=PUSH_PULL 2=PUSH_ONLY 3=PULL_ONLY
 *              4=PUSHPULL_FUSED (new: KVWorker::ZPushPull, one request + one reply per key)
 *   env :  DMLC_* topology, NUM_KEY_PER_SERVER (40), LOG_DURATION (10),
 *          TOTAL_DURATION, BENCHMARK_NTHREAD, ENABLE_RECV_BUFFER,
 *          TEST_NUM_GPU_WORKER / TEST_NUM_GPU_SERVER (>0: values live in HBM and
 *          travel over the NVLink van), SKIP_DEV_ID_CHECK, DEBUG_MODE
 * Output lines keep the reference's wording ("Application goodput: ... Gbps")
 * so existing log scrapers work; BENCH_JSON=1 adds one machine-readable line.
 * Differences: one GPU per process (PS_CUDA_DEVICE / LOCAL_RANK) instead of
 * key % local_size device hopping, and the store of a key is the buffer its
 * first push landed in (zero-copy) rather than a second allocation.
 */
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <thread>
#include <unordered_map>

#include "ps/ps.h"

#if PS_USE_CUDA
#include <cuda_runtime.h>
#define BENCH_CUDA(expr)                                                         \
  do {                                                                           \
    cudaError_t e_ = (expr);                                                     \
    CHECK(e_ == cudaSuccess) << "CUDA: " #expr ": " << cudaGetErrorString(e_);   \
  } while (0)
#endif

using namespace ps;

namespace {

enum Mode { PUSH_THEN_PULL = 0, PUSH_PULL = 1, PUSH_ONLY = 2, PULL_ONLY = 3, PUSHPULL_FUSED = 4 };

struct Options {
  int len = 1024000;
  int repeat = 10;
  Mode mode = PUSH_PULL;
  int keys_per_server = 40;
  int log_every = 10;
  int total_rounds = 2000000000;
  int nthread = 1;
  int num_ports = 1;
  bool recv_buffer = false;
  bool gpu_worker = false;
  bool gpu_server = false;
  bool skip_dev_check = false;
  bool debug = false;
  bool json = false;
  bool exportable = false;
  int group_size = 1;
};
Options opt;

int EnvInt(const char* k, int d) {
  const char* v = Environment::Get()->find(k);
  return v ? atoi(v) : d;
}

int MyDevice() {
#if PS_USE_CUDA
  int d = 0;
  if (cudaGetDevice(&d) == cudaSuccess) return d;
#endif
  return 0;
}

/*! \brief page-aligned host buffer or device buffer, wrapped with the right placement */
template <typename T>
SArray<T> AllocArray(size_t count, bool on_gpu, int dst_dev, bool dst_gpu, int src_dev = 0, int gpu_index = 0) {
  SArray<T> out;
  const size_t bytes = count * sizeof(T);
  if (on_gpu) {
#if PS_USE_CUDA
    // a process that drives several GPUs (DMLC_NUM_GPU_DEV) spreads its buffers over them
    const int first = MyDevice();
    if (gpu_index) BENCH_CUDA(cudaSetDevice(first + gpu_index));
    void* p = nullptr;
    BENCH_CUDA(cudaMalloc(&p, bytes));
    BENCH_CUDA(cudaMemset(p, 1, bytes));
    BENCH_CUDA(cudaDeviceSynchronize());
    out.reset(static_cast<T*>(p), count, [](T*) {}, GPU, first + gpu_index, dst_gpu ? GPU : CPU, dst_dev);
    if (gpu_index) BENCH_CUDA(cudaSetDevice(first));
#else
    LOG(FATAL) << "GPU buffers need a build with USE_CUDA=1";
#endif
  } else if (opt.exportable && sizeof(T) == 1) {
    // value buffers from the van's peer-mappable memory (shm arena on the shm van)
    void* p = Postoffice::Get()->van()->AllocExportable(bytes);
    CHECK(p);
    memset(p, 1, bytes);
    out.reset(static_cast<T*>(p), count, [](T*) {}, CPU, 0, dst_gpu ? GPU : CPU, dst_dev);
  } else {
    void* p = nullptr;
    const size_t page = static_cast<size_t>(sysconf(_SC_PAGESIZE));
    CHECK_EQ(posix_memalign(&p, page, (bytes + page - 1) / page * page), 0);
    memset(p, 1, bytes);
    out.reset(static_cast<T*>(p), count, [](T*) {}, CPU, src_dev, dst_gpu ? GPU : CPU, dst_dev);
  }
  return out;
}

struct KeySet {
  std::vector<SArray<Key>> keys;
  std::vector<SArray<char>> vals;
  std::vector<SArray<int>> lens;
};

/*! \brief key k lives on server k % S and is encoded as range[server].begin() + k */
KeySet MakeKeySet(int total_keys, bool vals_on_gpu, bool dst_gpu, int rank = 0) {
  KeySet ks;
  const auto& ranges = Postoffice::Get()->GetServerKeyRanges();
  const int S = static_cast<int>(ranges.size());
  for (int k = 0; k < total_keys; ++k) {
    SArray<Key> key = AllocArray<Key>(1, false, 0, false);
    key[0] = static_cast<Key>(ranges[k % S].begin() + k);
    SArray<int> len = AllocArray<int>(1, false, 0, false);
    len[0] = opt.len;
    // key k lives on device k % local_size at both ends (reference tests/test_benchmark.cc:58-90);
    // -1 lets a single-device receiver use the one it has
    const int ndev = EnvInt("PS_NUM_GPU_DEV", EnvInt("DMLC_NUM_GPU_DEV", 1));
    const int peer_first = EnvInt("TEST_PEER_GPU_BASE", -1);
    const int dst_dev = dst_gpu ? (ndev > 1 && peer_first >= 0 ? peer_first + (k / S) % ndev : -1) : (k % opt.num_ports);
    ks.keys.push_back(key);
    ks.lens.push_back(len);
    // multi-port vans pick the sending rail from the source context (reference src_key2ctx)
    ks.vals.push_back(AllocArray<char>(opt.len, vals_on_gpu, dst_dev, dst_gpu,
                                       (k + rank) % opt.num_ports, vals_on_gpu && ndev > 1 ? (k / S) % ndev : 0));
  }
  return ks;
}

// ---------------------------------------------------------------------------
// server
// ---------------------------------------------------------------------------
struct ServerState {
  std::mutex mu;
  std::unordered_map<uint64_t, KVPairs<char>> store;
  // (instance << 32 | worker id) -> decoded key -> registered buffer
  std::unordered_map<int64_t, std::unordered_map<uint64_t, SArray<char>>> registered;
};
ServerState g_server;

uint64_t DecodeKey(Key key) {
  return key - Postoffice::Get()->GetServerKeyRanges()[MyRank()].begin();
}

void ServerHandle(const KVMeta& req, const KVPairs<char>& data, KVServer<char>* server) {
  const uint64_t key = data.keys[0];
  if (!req.push) {
    KVPairs<char> res;
    {
      std::lock_guard<std::mutex> lk(g_server.mu);
      auto it = g_server.store.find(key);
      CHECK(it != g_server.store.end()) << "pull of a key that was never pushed: " << key;
      res = it->second;
    }
    server->Response(req, res);
    return;
  }
  CHECK(data.lens.size());
  CHECK_EQ(data.vals.size(), static_cast<size_t>(data.lens[0]))
      << "key=" << key << ", " << data.vals.size() << ", " << data.lens[0];
  const uint64_t decoded = DecodeKey(key);
  if (!opt.skip_dev_check && !opt.gpu_server) {
    CHECK_EQ(data.vals.dst_device_id_, static_cast<int>(decoded % opt.num_ports))
        << "key=" << decoded;
  }
  if (opt.gpu_server) CHECK(data.vals.on_gpu()) << "expected the push to land in HBM";
#if PS_USE_CUDA
  static const int ndev = EnvInt("PS_NUM_GPU_DEV", EnvInt("DMLC_NUM_GPU_DEV", 1));
  if (opt.gpu_server && ndev > 1 && EnvInt("TEST_CHECK_SLOT_DEVICE", 0) != 0) {
    // a server that drives several GPUs: key k must have landed on device first + (k / S) % ndev
    static const int first = EnvInt("PS_CUDA_DEVICE", 0);
    const int S = static_cast<int>(Postoffice::Get()->GetServerKeyRanges().size());
    cudaPointerAttributes attr;
    BENCH_CUDA(cudaPointerGetAttributes(&attr, data.vals.data()));
    CHECK_EQ(attr.device, first + static_cast<int>(decoded / S) % ndev)
        << "key " << decoded << " landed on the wrong device";
  }
#endif
  {
    std::lock_guard<std::mutex> lk(g_server.mu);
    if (!g_server.store.count(key)) {
      KVPairs<char>& slot = g_server.store[key];
      slot.keys.CopyFrom(data.keys);
      slot.lens.CopyFrom(data.lens);
      slot.vals = data.vals;  // the landing buffer doubles as the store
    }
    if (opt.recv_buffer) {
      const int64_t pair = (static_cast<int64_t>(server->instance_idx_) << 32) + req.sender;
      auto pit = g_server.registered.find(pair);
      CHECK(pit != g_server.registered.end()) << "no buffers registered for " << req.sender;
      auto bit = pit->second.find(decoded);
      CHECK(bit != pit->second.end()) << decoded;
      CHECK(bit->second.data() == data.vals.data())
          << "push did not land in the registered buffer: "
          << static_cast<const void*>(bit->second.data()) << " v.s. "
          << static_cast<const void*>(data.vals.data()) << " key=" << decoded
          << " sender=" << req.sender;
    }
  }
  if (opt.debug) {
    LOG(INFO) << "recved tensor! key=" << key << "\tlen: " << data.vals.size()
              << "\tsender: " << req.sender << "\taddr: " << static_cast<const void*>(data.vals.data());
  }
  if (req.pull) {  // fused push-pull: the stored values are the reply, there is no ack
    KVPairs<char> res;
    {
      std::lock_guard<std::mutex> lk(g_server.mu);
      res = g_server.store[key];
    }
    server->Response(req, res);
    return;
  }
  server->Response(req);  // empty ack
}

void StartServers(std::vector<KVServer<char>*>* servers) {
  if (!IsServer()) return;
  for (int i = 0; i < opt.group_size; ++i) {
    auto* s = new KVServer<char>(0, false, i);
    s->set_request_handle(ServerHandle);
    // neither the handler nor the (absent) worker callbacks wait for the network: both sides may
    // run on the receive threads (BENCHMARK_INLINE=0 restores the reference's thread structure)
    s->set_inline_dispatch(EnvInt("BENCHMARK_INLINE", 1) != 0);
    servers->push_back(s);
  }
  if (!opt.recv_buffer) return;
  const int W = NumWorkers(), S = NumServers();
  const int my_rank = MyRank();
  const int total_keys = S * opt.keys_per_server;
  for (int inst = 0; inst < opt.group_size; ++inst) {
    for (int w = 0; w < W; ++w) {
      KeySet ks = MakeKeySet(total_keys, opt.gpu_server, opt.gpu_server);
      const int worker_id = Postoffice::WorkerRankToID(w);
      for (int k = 0; k < total_keys; ++k) {
        if (k % S != my_rank) continue;
        (*servers)[inst]->RegisterRecvBufferWithRank(w, ks.keys[k], ks.vals[k], ks.lens[k]);
        const int64_t pair = (static_cast<int64_t>(inst) << 32) + worker_id;
        g_server.registered[pair][k] = ks.vals[k];
        KVPairs<char>& slot = g_server.store[ks.keys[k][0]];
        slot.keys = ks.keys[k];
        slot.vals = ks.vals[k];
        slot.lens = ks.lens[k];
      }
    }
  }
  Postoffice::Get()->Barrier(0, kWorkerGroup + kServerGroup);
}

// ---------------------------------------------------------------------------
// worker
// ---------------------------------------------------------------------------
void SteadyState(KVWorker<char>* kv, KeySet& ks, int total_keys, int tid) {
  const char* name = opt.mode == PUSH_PULL ? "PUSH_PULL"
                     : opt.mode == PUSH_ONLY ? "PUSH_ONLY"
                     : opt.mode == PULL_ONLY ? "PULL_ONLY" : "PUSHPULL_FUSED";
  LOG(INFO) << "========= " << name << " mode =========";
  LOG(INFO) << "========= msg_size=" << opt.len << " bytes =========";
  std::vector<int> in_flight;
  in_flight.reserve(static_cast<size_t>(total_keys) * 2);
  auto t0 = std::chrono::steady_clock::now();
  int rounds_in_window = 0;
  double best_gbps = 0;
  double issue_ns = 0, wait_ns = 0;  // per window: time spent issuing the round's requests / waiting for them
  Van* van = Postoffice::GetWorker(tid)->van();
  for (int round = 0; round < opt.total_rounds; ++round) {
    // all messages of a round are issued back to back: with PS_COALESCE_LAUNCHES their one-sided
    // copies share kernel launches and one completion event (a no-op otherwise)
    const auto t_round = std::chrono::steady_clock::now();
    {
      Van::CorkScope cork(van);  // released before the first Wait: the messages leave here
      for (int k = 0; k < total_keys; ++k) {
        if (opt.mode == PUSHPULL_FUSED) {
          in_flight.push_back(kv->ZPushPull(ks.keys[k], ks.vals[k], &ks.vals[k], &ks.lens[k]));
          continue;
        }
        if (opt.mode != PULL_ONLY) in_flight.push_back(kv->ZPush(ks.keys[k], ks.vals[k], ks.lens[k]));
        // the destination arrays must outlive the asynchronous pull (the reference
        // test passes pointers to loop locals, which only works by stack-slot luck)
        if (opt.mode != PUSH_ONLY) in_flight.push_back(kv->ZPull(ks.keys[k], &ks.vals[k], &ks.lens[k]));
      }
    }
    const auto t_issued = std::chrono::steady_clock::now();
    for (int ts : in_flight) kv->Wait(ts);
    in_flight.clear();
    issue_ns += std::chrono::duration<double, std::nano>(t_issued - t_round).count();
    wait_ns += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t_issued).count();
    if (++rounds_in_window % opt.log_every != 0) continue;
    const auto t1 = std::chrono::steady_clock::now();
    const double ns = std::chrono::duration<double, std::nano>(t1 - t0).count();
    const double gbps = 8.0 * opt.len * total_keys * rounds_in_window / ns;
    best_gbps = std::max(best_gbps, gbps);
    LL << "[" << tid << "]\tApplication goodput: " << gbps
       << " Gbps.\tAvg latency = " << ns / rounds_in_window / total_keys / 1000.0 << " ns per key"
       << " (issue " << issue_ns / rounds_in_window / total_keys / 1000.0 << " + wait "
       << wait_ns / rounds_in_window / total_keys / 1000.0 << " us per key)";
    issue_ns = wait_ns = 0;
    if (opt.json) {
      fprintf(stdout,
              "{\"bench\":\"test_benchmark\",\"mode\":%d,\"len\":%d,\"keys\":%d,\"tid\":%d,"
              "\"goodput_gbps\":%.4f,\"us_per_key\":%.3f}\n",
              static_cast<int>(opt.mode), opt.len, total_keys, tid, gbps,
              ns / rounds_in_window / total_keys / 1000.0);
      fflush(stdout);
    }
    rounds_in_window = 0;
    t0 = std::chrono::steady_clock::now();
  }
}

void RunWorker(KVWorker<char>* kv, int tid) {
  const int S = static_cast<int>(Postoffice::Get()->GetServerKeyRanges().size());
  CHECK_GT(S, 0);
  const int total_keys = S * opt.keys_per_server;
  KeySet ks = MakeKeySet(total_keys, opt.gpu_worker, opt.gpu_server, Postoffice::Get()->my_rank());
  if (opt.recv_buffer) {
    Postoffice::Get()->Barrier(0, kWorkerGroup + kServerGroup);
    LOG(INFO) << "Server recv buff registration is DONE.";
  }
  // warm-up / rendezvous: one acknowledged push per key, outside the timed region
  for (int k = 0; k < total_keys; ++k) kv->Wait(kv->ZPush(ks.keys[k], ks.vals[k], ks.lens[k]));

  if (opt.mode != PUSH_THEN_PULL) {
    SteadyState(kv, ks, total_keys, tid);
    return;
  }
  LOG(INFO) << "PUSH_THEN_PULL mode";
  double push_ns = 0, pull_ns = 0;
  for (int r = 0; r < opt.repeat; ++r) {
    auto a = std::chrono::steady_clock::now();
    for (int s = 0; s < S; ++s) kv->Wait(kv->ZPush(ks.keys[s], ks.vals[s], ks.lens[s]));
    push_ns += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - a).count();
  }
  LL << "push " << opt.len << " bytes to each server, repeat=" << opt.repeat
     << ", total_time=" << push_ns / 1e6 << "ms";
  for (int r = 0; r < opt.repeat; ++r) {
    auto a = std::chrono::steady_clock::now();
    for (int s = 0; s < S; ++s) kv->Wait(kv->ZPull(ks.keys[s], &ks.vals[s], &ks.lens[s]));
    pull_ns += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - a).count();
  }
  LL << "pull " << opt.len << " bytes to each server, repeat=" << opt.repeat
     << ", total_time=" << pull_ns / 1e6 << "ms";
}

}  // namespace

int main(int argc, char* argv[]) {
  opt.len = argc > 1 ? atoi(argv[1]) : 1024000;
  opt.repeat = argc > 2 ? atoi(argv[2]) : 10;
  opt.mode = argc > 3 ? static_cast<Mode>(atoi(argv[3])) : PUSH_PULL;
  opt.keys_per_server = EnvInt("NUM_KEY_PER_SERVER", 40);
  opt.log_every = std::max(1, EnvInt("LOG_DURATION", 10));
  opt.total_rounds = EnvInt("TOTAL_DURATION", 2000000000);
  opt.nthread = EnvInt("BENCHMARK_NTHREAD", 1);
  opt.num_ports = std::max(1, EnvInt("DMLC_NUM_PORTS", 1));
  opt.recv_buffer = EnvInt("ENABLE_RECV_BUFFER", 0) != 0;
  opt.gpu_worker = EnvInt("TEST_NUM_GPU_WORKER", 0) > 0;
  opt.gpu_server = EnvInt("TEST_NUM_GPU_SERVER", 0) > 0;
  opt.skip_dev_check = EnvInt("SKIP_DEV_ID_CHECK", 0) != 0;
  opt.debug = Environment::Get()->find("DEBUG_MODE") != nullptr;
  opt.json = EnvInt("BENCH_JSON", 0) != 0;
  opt.exportable = EnvInt("TEST_EXPORTABLE_VALS", 0) != 0 ||
                   std::string(GetEnv("PS_VAN_TYPE", "")) == "shm";
  opt.group_size = std::max(1, EnvInt("DMLC_GROUP_SIZE", 1));
  LOG(INFO) << opt.num_ports << " ports per node; recv buffer registration is "
            << (opt.recv_buffer ? "enabled" : "NOT enabled") << "; gpu worker/server = "
            << opt.gpu_worker << "/" << opt.gpu_server;

  const std::string role_str = CHECK_NOTNULL(Environment::Get()->find("DMLC_ROLE"));
  const Node::Role role = GetRole(role_str);
#if PS_USE_CUDA
  if (opt.gpu_worker || opt.gpu_server) {
    int ndev = 0;
    BENCH_CUDA(cudaGetDeviceCount(&ndev));
    int dev = EnvInt("PS_CUDA_DEVICE", EnvInt("LOCAL_RANK", 0)) % std::max(1, ndev);
    BENCH_CUDA(cudaSetDevice(dev));
  }
#endif
  const int my_rank = EnvInt("DMLC_RANK", -1);
  StartPS(0, role, my_rank, true);
  if (my_rank != -1 && role != Node::SCHEDULER) {
    CHECK_EQ(MyRank(), my_rank) << "scheduler ignored the preferred rank";
  }

  std::vector<KVServer<char>*> servers;
  StartServers(&servers);
  if (!IsServer() && !IsScheduler()) {
    LOG(INFO) << "number of threads for the same worker = " << opt.nthread;
    std::vector<KVWorker<char>*> kvs;
    std::vector<std::thread> threads;
    for (int i = 0; i < opt.nthread; ++i) {
      kvs.push_back(new KVWorker<char>(0, 0, i));
      kvs.back()->set_inline_dispatch(EnvInt("BENCHMARK_INLINE", 1) != 0);
      threads.emplace_back(RunWorker, kvs.back(), i);
    }
    for (auto& t : threads) t.join();
  }
  Finalize(0, role, true);
  for (auto* s : servers) delete s;
  return 0;
}
