// In-process cluster smoke test: scheduler + 1 server + 1 worker as threads.
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
  std::thread fs([] { Postoffice::GetScheduler()->Finalize(0, true); });
  std::thread fv([] { Postoffice::GetServer()->Finalize(0, true); });
  std::thread fw([] { Postoffice::GetWorker()->Finalize(0, true); });
  fs.join(); fv.join(); fw.join();
  delete server;
  LOG(INFO) << "done";
  return 0;
}
