/**
 * \file test_simple_app.cc
 * \brief SimpleApp request/response between every pair of roles (historical
 * test_simple_app of ps-lite; absent from the reference fork).
 * Workers send `n` requests to the server group and one to the scheduler; handlers
 * echo the head and a body; the worker checks counts and payloads.
 */
#include "ps/ps.h"
using namespace ps;

int main(int argc, char* argv[]) {
  const int n = argc > 1 ? atoi(argv[1]) : 50;
  const std::string role_str = CHECK_NOTNULL(Environment::Get()->find("DMLC_ROLE"));
  const Node::Role role = GetRole(role_str);
  StartPS(0, role, -1, true);
  std::atomic<int> requests{0}, responses{0}, bad{0};
  SimpleApp app(0, 0);
  app.set_request_handle([&](const SimpleData& req, SimpleApp* a) {
    ++requests;
    a->Response(req, "echo:" + req.body);
  });
  app.set_response_handle([&](const SimpleData& res, SimpleApp*) {
    ++responses;
    if (res.body.rfind("echo:", 0) != 0) ++bad;
  });
  // everyone's handlers are installed before any request flies
  Postoffice::Get()->Barrier(0, kWorkerGroup + kServerGroup + kScheduler);
  int rc = 0;
  if (role == Node::WORKER) {
    std::vector<int> ts;
    for (int i = 0; i < n; ++i) ts.push_back(app.Request(i, "msg" + std::to_string(i), kServerGroup));
    ts.push_back(app.Request(1000, "to-scheduler", kScheduler));
    for (int t : ts) app.Wait(t);
    const int expect = n * NumServers() + 1;
    if (responses.load() != expect || bad.load()) {
      LL << "expected " << expect << " good responses, got " << responses.load() << " (bad " << bad.load() << ")";
      rc = 1;
    }
    LL << (rc ? "test_simple_app FAILED" : "test_simple_app PASSED");
  }
  Postoffice::Get()->Barrier(0, kWorkerGroup + kServerGroup + kScheduler);
  if (role == Node::SERVER && requests.load() != n * NumWorkers()) {
    LL << "server saw " << requests.load() << " requests, expected " << n * NumWorkers();
    rc = 1;
  }
  Finalize(0, role, true);
  return rc;
}
