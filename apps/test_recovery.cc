/**
 * \file test_recovery.cc
 * \brief Failure detection + recovery re-registration.
 *
 * Run with PS_HEARTBEAT_INTERVAL=1 PS_HEARTBEAT_TIMEOUT=2. Scenario driven by env:
 *   server, scheduler      : normal life cycle, the server sums pushes per key
 *   worker, RECOVERY_CRASH=1   : pushes once, then dies without Finalize (_exit)
 *   worker, (late starter)     : registers after the cluster is complete; the scheduler
 *                                hands it the dead worker's id; it must see is_recovery(),
 *                                reach the server again and read the state the dead node left
 * The surviving processes skip the final barrier partner that died by finalising without a
 * barrier (as the reference prescribes for recovery nodes, postoffice.h:230).
 */
#include <unistd.h>
#include <chrono>
#include <thread>
#include "ps/ps.h"
using namespace ps;

int main() {
  const std::string role_str = CHECK_NOTNULL(Environment::Get()->find("DMLC_ROLE"));
  const Node::Role role = GetRole(role_str);
  const bool crash = GetEnv("RECOVERY_CRASH", 0) != 0;
  const bool late = GetEnv("RECOVERY_LATE", 0) != 0;
  // a recovery node cannot take part in the start-up barrier (everyone else passed it long ago)
  StartPS(0, role, -1, !late);
  if (role == Node::SCHEDULER) {
    // stay up until the replacement has registered and finished; it tells us via SimpleApp
    SimpleApp app(0, 0);
    std::atomic<int> done{0};
    app.set_request_handle([&](const SimpleData& req, SimpleApp* a) {
      ++done;
      a->Response(req);
    });
    while (done.load() < 2) std::this_thread::sleep_for(std::chrono::milliseconds(50));
    auto dead = Postoffice::Get()->GetDeadNodes(2);
    LL << "scheduler: dead nodes now = " << dead.size();
    Finalize(0, role, false);
    return 0;
  }
  if (role == Node::SERVER) {
    KVServer<float> server(0);
    server.set_request_handle(KVServerDefaultHandle<float>());
    std::atomic<int> done{0};
    SimpleApp app(1, 1);
    app.set_request_handle([&](const SimpleData& req, SimpleApp* a) {
      ++done;
      a->Response(req);
    });
    while (done.load() < 1) std::this_thread::sleep_for(std::chrono::milliseconds(50));
    Finalize(0, role, false);
    return 0;
  }
  KVWorker<float> kv(0, 0);
  std::vector<Key> keys = {7};
  // RECOVERY_ONESIDED=1 (one-sided vans): values live in exportable memory, so pushes and pull replies
  // are one-sided writes into regions the peers have mapped. The replacement process has the dead
  // worker's node id but its own regions: everything the survivors cached about that id must go.
  const bool onesided = GetEnv("RECOVERY_ONESIDED", 0) != 0;
  Van* van = Postoffice::Get()->van();
  auto exportable = [&](float v) {
    float* p = static_cast<float*>(van->AllocExportable(256));
    CHECK(p) << "RECOVERY_ONESIDED needs a van with exportable memory (shm / nvl)";
    p[0] = v;
    SArray<float> a;
    a.reset(p, 1, [](float*) {});
    return a;
  };
  if (onesided) {
    SArray<Key> zkeys(keys);
    if (crash) {
      SArray<float> vals = exportable(5.f), got = exportable(-1.f);
      kv.Wait(kv.ZPush(zkeys, vals));
      kv.Wait(kv.ZPull(zkeys, &got));  // the server now has this process's region mapped
      CHECK_GE(got[0], 5.f);
      LL << "worker " << MyRank() << " crashing after its one-sided push and pull";
      _exit(0);
    }
    if (late) {
      CHECK(Postoffice::Get()->is_recovery()) << "late worker was not flagged as a recovery node";
      LL << "recovery worker adopted rank " << MyRank() << " id " << van->my_node().id;
      SArray<float> got = exportable(-1.f);
      kv.Wait(kv.ZPull(zkeys, &got));
      CHECK_EQ(got[0], 6.f) << "the pull reply did not land in the replacement's memory";
      SArray<float> vals = exportable(10.f);
      kv.Wait(kv.ZPush(zkeys, vals));
      kv.Wait(kv.ZPull(zkeys, &got));
      CHECK_EQ(got[0], 16.f);
      LL << "test_recovery PASSED";
      SimpleApp app(1, 2);
      app.Wait(app.Request(0, "done", kServerGroup));
      SimpleApp app2(0, 3);
      app2.Wait(app2.Request(0, "done", kScheduler));
      Finalize(0, role, false);
      return 0;
    }
    SArray<float> vals = exportable(1.f), got = exportable(-1.f);
    kv.Wait(kv.ZPush(zkeys, vals));
    kv.Wait(kv.ZPull(zkeys, &got));
    SimpleApp app2(0, 3);
    std::this_thread::sleep_for(std::chrono::seconds(GetEnv("RECOVERY_SURVIVOR_WAIT", 8)));
    app2.Wait(app2.Request(0, "done", kScheduler));
    Finalize(0, role, false);
    return 0;
  }
  if (crash) {
    std::vector<float> vals = {5.f};
    kv.Wait(kv.Push(keys, vals));
    LL << "worker " << MyRank() << " crashing after its push";
    _exit(0);
  }
  if (late) {
    CHECK(Postoffice::Get()->is_recovery()) << "late worker was not flagged as a recovery node";
    LL << "recovery worker adopted rank " << MyRank() << " id " << Postoffice::Get()->van()->my_node().id;
    std::vector<float> got;
    kv.Wait(kv.Pull(keys, &got));
    CHECK_EQ(got.size(), (size_t)1);
    // 5 from the dead worker + 1 from the survivor
    CHECK_EQ(got[0], 6.f) << "state left by the dead worker is not visible";
    std::vector<float> vals = {10.f};
    kv.Wait(kv.Push(keys, vals));
    kv.Wait(kv.Pull(keys, &got));
    CHECK_EQ(got[0], 16.f);
    LL << "test_recovery PASSED";
    // release the scheduler and the server
    SimpleApp app(1, 2);
    app.Wait(app.Request(0, "done", kServerGroup));
    SimpleApp app2(0, 3);
    app2.Wait(app2.Request(0, "done", kScheduler));
    Finalize(0, role, false);
    return 0;
  }
  // the surviving worker
  std::vector<float> vals = {1.f};
  kv.Wait(kv.Push(keys, vals));
  SimpleApp app2(0, 3);
  // wait for the replacement to show up before leaving, then tell the scheduler
  std::this_thread::sleep_for(std::chrono::seconds(GetEnv("RECOVERY_SURVIVOR_WAIT", 8)));
  app2.Wait(app2.Request(0, "done", kScheduler));
  Finalize(0, role, false);
  return 0;
}
