// The C++ example of docs/tutorials.md §1, kept compilable:
//   g++ -std=c++17 -Iinclude -Isrc examples/kv_hello.cc build/libpslite.a -pthread -lrt ... -o kv_hello
//   scripts/local.sh 2 2 build/kv_hello
#include <cmath>
#include "ps/ps.h"
using namespace ps;

int main() {
  const Node::Role role = GetRole(CHECK_NOTNULL(Environment::Get()->find("DMLC_ROLE")));
  StartPS(0, role, -1, true);  // connect + barrier
  KVServer<float>* server = nullptr;
  if (IsServer()) {
    server = new KVServer<float>(0);
    server->set_request_handle(KVServerDefaultHandle<float>());  // store[key] += val
  }
  int bad = 0;
  if (role == Node::WORKER) {
    KVWorker<float> kv(0, 0);
    std::vector<Key> keys = {1, 3, 5};
    std::vector<float> vals = {1.f, 2.f, 3.f}, got;
    kv.Wait(kv.Push(keys, vals));
    Postoffice::GetWorker()->Barrier(0, kWorkerGroup);  // every worker has pushed
    kv.Wait(kv.Pull(keys, &got));                       // got == sum over all workers
    for (size_t i = 0; i < vals.size(); ++i) {
      if (std::fabs(got[i] - NumWorkers() * vals[i]) > 1e-5) ++bad;
    }
    LOG(INFO) << "kv_hello " << (bad ? "FAILED" : "PASSED") << ": " << got[0] << " " << got[1] << " " << got[2];
  }
  Finalize(0, role, true);
  delete server;
  return bad;
}
