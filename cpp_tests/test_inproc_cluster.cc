// In-process cluster smoke test: scheduler + 1 server + 1 worker as threads.
#include <map>
#include <set>
#include "ps/ps.h"
using namespace ps;
int main() {
  Environment::Init({{"DMLC_NUM_WORKER", "1"}, {"DMLC_NUM_SERVER", "1"},
                     {"DMLC_PS_ROOT_URI", "127.0.0.1"}, {"DMLC_PS_ROOT_PORT", "18777"},
                     {"DMLC_NODE_HOST", "127.0.0.1"}});
  Postoffice::Init(Node::SCHEDULER);
  Postoffice::Init(Node::JOINT);
  std::thread ts([] { Postoffice::GetScheduler()->Start(0, Node::SCHEDULER, -1, true, nullptr); });
  std::thread tv([] { Postoffice::GetServer()->Start(0, Node::SERVER, -1, true, nullptr); });
  std::thread tw([] { Postoffice::GetWorker()->Start(0, Node::WORKER, -1, true, nullptr); });
  ts.join(); tv.join(); tw.join();
  LOG(INFO) << "cluster up";
  // node numbering of the reference: scheduler 1, server rank r -> 2r+8, worker rank r -> 2r+9;
  // group ids are bit masks whose members are the union of the named groups
  CHECK_EQ(Postoffice::ServerRankToID(0), 8); CHECK_EQ(Postoffice::WorkerRankToID(0), 9);
  CHECK_EQ(Postoffice::ServerRankToID(3), 14); CHECK_EQ(Postoffice::WorkerRankToID(3), 15);
  CHECK_EQ(Postoffice::IDtoRank(14), 3); CHECK_EQ(Postoffice::IDtoRank(15), 3);
  {
    Postoffice* po = Postoffice::GetWorker();
    CHECK_EQ(po->GetNodeIDs(kScheduler).size(), (size_t)1);
    CHECK_EQ(po->GetNodeIDs(kScheduler)[0], 1);
    CHECK_EQ(po->GetNodeIDs(kServerGroup).size(), (size_t)1);
    CHECK_EQ(po->GetNodeIDs(kServerGroup)[0], 8);
    CHECK_EQ(po->GetNodeIDs(kWorkerGroup)[0], 9);
    CHECK_EQ(po->GetNodeIDs(kWorkerGroup + kServerGroup).size(), (size_t)2);
    CHECK_EQ(po->GetNodeIDs(kWorkerGroup + kServerGroup + kScheduler).size(), (size_t)3);
    CHECK_EQ(po->num_workers(), 1); CHECK_EQ(po->num_servers(), 1);
    CHECK(po->is_worker()); CHECK(!po->is_server());
    CHECK_EQ(po->my_rank(), 0);
    const auto& ranges = po->GetServerKeyRanges();
    CHECK_EQ(ranges.size(), (size_t)1);
    CHECK_EQ(ranges[0].begin(), (uint64_t)0);
    CHECK(Postoffice::GetServer()->is_server());
    CHECK(Postoffice::GetScheduler()->is_scheduler());
  }
  auto* server = new KVServer<float>(0);
  server->set_request_handle(KVServerDefaultHandle<float>());
  KVWorker<float> kv(0, 0);
  std::vector<Key> keys = {1, 3, 5};
  std::vector<float> vals = {1.5f, 2.5f, 3.5f};
  kv.Wait(kv.Push(keys, vals));
  kv.Wait(kv.Push(keys, vals));
  std::vector<float> out;
  kv.Wait(kv.Pull(keys, &out));
  CHECK_EQ(out.size(), (size_t)3);
  CHECK_EQ(out[0], 3.0f); CHECK_EQ(out[2], 7.0f);
  LOG(INFO) << "push/pull ok: " << out[0] << " " << out[1] << " " << out[2];

  // edges of the key space: the first key, the last key below kMaxKey, an isolated key
  std::vector<Key> edge = {0, 7, kMaxKey - 1};
  std::vector<float> ev = {10.f, 20.f, 30.f};
  kv.Wait(kv.Push(edge, ev));
  std::vector<float> eo;
  kv.Wait(kv.Pull(edge, &eo));
  CHECK_EQ(eo.size(), (size_t)3);
  CHECK_EQ(eo[0], 10.f); CHECK_EQ(eo[1], 20.f); CHECK_EQ(eo[2], 30.f);

  // pulling into a larger pre-sized vector is allowed: only the front is written
  std::vector<float> big(8, -1.f);
  kv.Wait(kv.Pull(keys, &big));
  CHECK_EQ(big.size(), (size_t)8);
  CHECK_EQ(big[0], 3.0f); CHECK_EQ(big[2], 7.0f); CHECK_EQ(big[3], -1.f);

  // callbacks run exactly once, before Wait returns
  std::atomic<int> fired{0};
  kv.Wait(kv.Push(keys, vals, {}, 0, [&fired] { ++fired; }));
  CHECK_EQ(fired.load(), 1);

  // fused push-pull: push once more and get the new sums in one round trip
  std::vector<float> pp;
  kv.Wait(kv.PushPull(keys, vals, &pp));
  CHECK_EQ(pp.size(), (size_t)3);
  CHECK_EQ(pp[0], 1.5f * 4); CHECK_EQ(pp[2], 3.5f * 4);

  // zero-copy variants keep working on caller-owned arrays
  SArray<Key> zk(keys);
  SArray<float> zv(3, 0.f);
  kv.Wait(kv.ZPull(zk, &zv));
  CHECK_EQ(zv[1], 2.5f * 4);
  SArray<float> zin(vals), zout(3, 0.f);
  kv.Wait(kv.ZPushPull(zk, zin, &zout));
  CHECK_EQ(zout[1], 2.5f * 5);
  // variable-length values (lens): a second app with a handler that stores a vector per key
  auto* vserver = new KVServer<float>(1);
  auto vstore = std::make_shared<std::map<Key, std::vector<float>>>();
  vserver->set_request_handle([vstore](const KVMeta& req, const KVPairs<float>& d, KVServer<float>* srv) {
    KVPairs<float> res;
    if (req.push) {
      CHECK_EQ(d.lens.size(), d.keys.size());
      size_t at = 0;
      for (size_t i = 0; i < d.keys.size(); ++i) {
        (*vstore)[d.keys[i]].assign(d.vals.data() + at, d.vals.data() + at + d.lens[i]);
        at += static_cast<size_t>(d.lens[i]);
      }
      CHECK_EQ(at, d.vals.size());
    } else {
      res.keys = d.keys;
      for (Key k : d.keys) {
        const auto& v = (*vstore)[k];
        res.lens.push_back(static_cast<int>(v.size()));
        for (float x : v) res.vals.push_back(x);
      }
    }
    srv->Response(req, res);
  });
  {
    KVWorker<float> vkv(1, 1);
    std::vector<Key> vk = {2, 9, 11};
    std::vector<float> vv = {1.f, 2.f, 2.5f, 3.f, 3.25f, 3.5f};
    std::vector<int> vl = {1, 2, 3};
    vkv.Wait(vkv.Push(vk, vv, vl));
    std::vector<float> got;   // 0-sized: filled by the pull
    std::vector<int> got_len; // 0-sized: filled by the pull
    vkv.Wait(vkv.Pull(vk, &got, &got_len));
    CHECK_EQ(got.size(), (size_t)6);
    CHECK_EQ(got_len.size(), (size_t)3);
    CHECK_EQ(got_len[0], 1); CHECK_EQ(got_len[2], 3);
    CHECK_EQ(got[1], 2.f); CHECK_EQ(got[5], 3.5f);
  }
  delete vserver;

  // inline dispatch: handlers run on the receive thread while the customer queue is idle and on
  // the customer thread otherwise; either way one at a time and in arrival order. A second app
  // whose handler checks both, fed with bursts of back-to-back requests so that both paths occur
  {
    auto* iserver = new KVServer<float>(11);
    std::atomic<int> in_handler{0}, handled{0}, out_of_order{0}, overlapped{0};
    std::set<std::thread::id> handler_threads;
    float last = 0.f;
    iserver->set_request_handle([&](const KVMeta& req, const KVPairs<float>& data, KVServer<float>* srv) {
      if (in_handler.fetch_add(1) != 0) ++overlapped;
      handler_threads.insert(std::this_thread::get_id());
      if (req.push) {
        if (data.vals[0] != last + 1.f) ++out_of_order;
        last = data.vals[0];
      }
      ++handled;
      in_handler.fetch_sub(1);
      srv->Response(req);
    });
    iserver->set_inline_dispatch(true);
    KVWorker<float> ikv(11, 1);
    ikv.set_inline_dispatch(true);
    std::vector<Key> one = {4};
    float next = 1.f;
    for (int burst = 0; burst < 50; ++burst) {
      std::vector<int> ts;
      // every other burst starts queued and switches to inline half-way: the switch must not let
      // an inline message overtake the queued ones or run beside them
      if (burst % 2 == 1) iserver->set_inline_dispatch(false);
      for (int i = 0; i < 40; ++i, next += 1.f) {
        if (i == 20) iserver->set_inline_dispatch(true);
        ts.push_back(ikv.Push(one, std::vector<float>{next}));
      }
      for (int t : ts) ikv.Wait(t);
      if (burst % 10 == 9) std::this_thread::sleep_for(std::chrono::milliseconds(2));  // let the queue drain
    }
    CHECK_EQ(handled.load(), 2000);
    CHECK_EQ(out_of_order.load(), 0);
    CHECK_EQ(overlapped.load(), 0);
    LOG(INFO) << "inline dispatch ok: 2000 requests on " << handler_threads.size() << " handler thread(s)";
    delete iserver;
  }

  // SimpleApp surface of the KV classes: a request to the server group, answered with a body
  server->SimpleApp::set_request_handle([](const SimpleData& req, SimpleApp* app) {
    app->Response(req, "echo:" + req.body);
  });
  std::string reply;
  kv.SimpleApp::set_response_handle([&reply](const SimpleData& res, SimpleApp*) { reply = res.body; });
  kv.SimpleApp::Wait(kv.Request(42, "ping", kServerGroup));
  CHECK_EQ(reply, std::string("echo:ping"));
  LOG(INFO) << "edge cases ok";
  std::atomic<int> exit_calls{0};
  Postoffice::GetWorker()->RegisterExitCallback([&exit_calls] { ++exit_calls; });
  std::thread fs([] { Postoffice::GetScheduler()->Finalize(0, true); });
  std::thread fv([] { Postoffice::GetServer()->Finalize(0, true); });
  std::thread fw([] { Postoffice::GetWorker()->Finalize(0, true); });
  fs.join(); fv.join(); fw.join();
  CHECK_EQ(exit_calls.load(), 1);  // the exit callback ran during the worker's Finalize
  delete server;
  LOG(INFO) << "done";
  return 0;
}
