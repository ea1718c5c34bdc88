/**
 * \file test_connection.cc
 * \brief Smallest possible job: StartPS, a few barriers, Finalize (historical
 * test_connection). Also checks rank assignment and the id maps.
 */
#include "ps/ps.h"
using namespace ps;
int main(int, char*[]) {
  const std::string role_str = CHECK_NOTNULL(Environment::Get()->find("DMLC_ROLE"));
  const Node::Role role = GetRole(role_str);
  const int want = GetEnv("DMLC_RANK", -1);
  StartPS(0, role, want, true);
  if (role != Node::SCHEDULER) {
    CHECK_GE(MyRank(), 0);
    if (want >= 0) CHECK_EQ(MyRank(), want);
    const int id = Postoffice::Get()->van()->my_node().id;
    CHECK_EQ(Postoffice::IDtoRank(id), Postoffice::Get()->my_rank());
    CHECK_EQ(id % 2, IsServer() ? 0 : 1);
  }
  for (int i = 0; i < 3; ++i) Postoffice::Get()->Barrier(0, kWorkerGroup + kServerGroup + kScheduler);
  if (role == Node::WORKER) Postoffice::Get()->Barrier(0, kWorkerGroup);
  if (role == Node::SERVER) Postoffice::Get()->Barrier(0, kServerGroup);
  Finalize(0, role, true);
  LL << "test_connection PASSED (" << role_str << ")";
  return 0;
}
