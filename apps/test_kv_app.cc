/**
 * \file test_kv_app.cc
 * \brief Functional test of KVWorker / KVServer across real processes.
 *
 * Every worker pushes `repeat` times a multi-key list whose keys span *all* server key
 * ranges (so the slicer, per-server fan-out and response counting are exercised), then
 * all workers barrier and pull: the result must equal repeat * num_workers * vals.
 * Works for every role including `joint` (worker + server in one process) and with
 * DMLC_GROUP_SIZE > 1 (one KVWorker / KVServer per instance). The reference ships no such
 * test any more (its travis script still names test_kv_app, tests/travis/travis_script.sh).
 *   usage: test_kv_app [num_keys=1000] [vals_per_key=3] [repeat=5]
 */
#include <cmath>
#include "ps/ps.h"
using namespace ps;

int main(int argc, char* argv[]) {
  const int num_keys = argc > 1 ? atoi(argv[1]) : 1000;
  const int k = argc > 2 ? atoi(argv[2]) : 3;
  const int repeat = argc > 3 ? atoi(argv[3]) : 5;
  const std::string role_str = CHECK_NOTNULL(Environment::Get()->find("DMLC_ROLE"));
  const Node::Role role = GetRole(role_str);
  const int group_size = std::max(1, GetEnv("DMLC_GROUP_SIZE", 1));
  const int rank = GetEnv("DMLC_RANK", group_size > 1 ? 0 : -1);
  StartPS(0, role, rank, true);

  const bool has_server = role == Node::SERVER || role == Node::JOINT;
  const bool has_worker = role == Node::WORKER || role == Node::JOINT;
  std::vector<KVServer<float>*> servers;
  if (has_server) {
    for (int i = 0; i < group_size; ++i) {
      auto* s = new KVServer<float>(0, false, i);
      // one scalar store per server instance; values are k floats per key -> store per element
      auto store = std::make_shared<std::unordered_map<Key, std::vector<float>>>();
      s->set_request_handle([store, k](const KVMeta& req, const KVPairs<float>& data, KVServer<float>* srv) {
        const size_t n = data.keys.size();
        KVPairs<float> res;
        if (req.push) {
          CHECK_EQ(data.vals.size(), n * static_cast<size_t>(k));
          for (size_t i = 0; i < n; ++i) {
            auto& v = (*store)[data.keys[i]];
            v.resize(k, 0.f);
            for (int j = 0; j < k; ++j) v[j] += data.vals[i * k + j];
          }
        }
        if (!req.push || req.pull) {  // a pull, or the reply half of a fused push-pull
          res.keys = data.keys;
          res.vals.resize(n * k);
          for (size_t i = 0; i < n; ++i) {
            auto& v = (*store)[data.keys[i]];
            v.resize(k, 0.f);
            for (int j = 0; j < k; ++j) res.vals[i * k + j] = v[j];
          }
        }
        srv->Response(req, res);
      });
      servers.push_back(s);
    }
  }
  int failures = 0;
  if (has_worker) {
    std::vector<std::thread> threads;
    std::atomic<int> bad{0};
    for (int inst = 0; inst < group_size; ++inst) {
      threads.emplace_back([&, inst] {
        KVWorker<float> kv(0, 0, inst);
        std::vector<Key> keys(num_keys);
        std::vector<float> vals(static_cast<size_t>(num_keys) * k);
        const Key stride = kMaxKey / static_cast<Key>(num_keys);
        for (int i = 0; i < num_keys; ++i) {
          keys[i] = stride * static_cast<Key>(i) + static_cast<Key>(inst);
          for (int j = 0; j < k; ++j) vals[i * k + j] = 0.5f * (i % 97) + j;
        }
        std::vector<int> ts;
        for (int r = 0; r < repeat; ++r) ts.push_back(kv.Push(keys, vals));
        for (int t : ts) kv.Wait(t);
        if (inst == 0) Postoffice::GetWorker(0)->Barrier(0, kWorkerGroup);
        else std::this_thread::sleep_for(std::chrono::milliseconds(300));
        std::vector<float> got;
        kv.Wait(kv.Pull(keys, &got));
        const float scale = static_cast<float>(repeat * Postoffice::GetWorker(inst)->num_workers());
        double err = 0;
        for (size_t i = 0; i < vals.size(); ++i) err += std::fabs(got[i] - scale * vals[i]);
        if (err > 1e-3 * vals.size()) {
          ++bad;
          LL << "instance " << inst << " mismatch, total abs error " << err;
        }
        // zero-copy variants with a callback
        SArray<Key> zk(keys);
        SArray<float> zv(vals.size(), 0.f);
        std::atomic<int> fired{0};
        kv.Wait(kv.ZPull(zk, &zv, nullptr, 0, [&fired] { ++fired; }));
        if (fired.load() != 1 || std::fabs(zv[1] - scale * vals[1]) > 1e-3) ++bad;
        // host one-sided van: a multi-key pull into exportable (shared) memory lands in place —
        // every server writes its slice of the list straight into this buffer
        Van* van = Postoffice::GetWorker(inst)->van();
        if (van->GetType() == "shm") {
          const size_t bytes = vals.size() * sizeof(float);
          float* raw = static_cast<float*>(van->AllocExportable(bytes));
          memset(raw, 0, bytes);
          SArray<float> ev(raw, vals.size(), false);
          kv.Wait(kv.ZPull(zk, &ev));
          double e2 = 0;
          for (size_t i = 0; i < vals.size(); ++i) e2 += std::fabs(raw[i] - scale * vals[i]);
          if (e2 > 1e-3 * vals.size()) {
            ++bad;
            LL << "instance " << inst << " exportable multi-key pull mismatch " << e2;
          }
          van->FreeExportable(raw);
        }
        // every worker has finished reading the sums above before anyone changes them again
        if (inst == 0) Postoffice::GetWorker(0)->Barrier(0, kWorkerGroup);
        else std::this_thread::sleep_for(std::chrono::milliseconds(300));
        // fused push-pull: one more contribution from every worker, one message pair per server.
        // Workers race, so a worker sees between 1 and W of the new contributions.
        {
          std::vector<float> after;
          kv.Wait(kv.PushPull(keys, vals, &after));
          const float W = static_cast<float>(Postoffice::GetWorker(inst)->num_workers());
          int off_grid = 0;
          for (size_t i = 0; i < vals.size(); ++i) {
            if (vals[i] == 0.f) continue;
            const float extra = after[i] / vals[i] - scale;  // how many new pushes it already contains
            if (extra < 1.f - 1e-3f || extra > W + 1e-3f || std::fabs(extra - std::round(extra)) > 1e-3f) ++off_grid;
          }
          if (after.size() != vals.size() || off_grid) {
            ++bad;
            LL << "instance " << inst << " push-pull: " << off_grid << " values off the expected grid";
          }
        }
      });
    }
    for (auto& t : threads) t.join();
    failures = bad.load();
    LL << (failures ? "test_kv_app FAILED" : "test_kv_app PASSED") << " on worker rank " << MyRank();
  }
  Finalize(0, role, true);
  for (auto* s : servers) delete s;
  return failures ? 1 : 0;
}
