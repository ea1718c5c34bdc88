/**
 * \file ps.h
 * \brief Umbrella header: start / stop the system and query the topology.
 *
 * API parity: reference include/ps/ps.h:16-30 (NumWorkers ... MyRank), :38-52
 * (GetRole), :110-138 (StartPS), :183-192 (Finalize), :209-211
 * (RegisterExitCallback). Roles: worker, server, scheduler, and joint (one
 * worker + one server instance in the same process — on B200 the co-located
 * pair shares one GPU's HBM, so a push between them is a pointer hand-off).
 * DMLC_GROUP_SIZE=N starts N instances per role in the process (one per rail /
 * device); instance i of a worker group only talks to instance i of each
 * server group.
 */
#ifndef PS_PS_H_
#define PS_PS_H_
#include <functional>
#include <string>
#include <thread>
#include <vector>
#include "ps/base.h"
#include "ps/kv_app.h"
#include "ps/simple_app.h"

namespace ps {

/*! \brief number of worker groups */
inline int NumWorkers() { return Postoffice::Get()->num_workers(); }
/*! \brief number of server groups */
inline int NumServers() { return Postoffice::Get()->num_servers(); }
inline bool IsServer() { return Postoffice::Get()->is_server(); }
inline bool IsScheduler() { return Postoffice::Get()->is_scheduler(); }
/*! \brief group-level rank of this node within its role; valid after StartPS */
inline int MyRank() { return Postoffice::Get()->my_rank() / Postoffice::Get()->group_size(); }

/*! \brief "worker" | "server" | "scheduler" | "joint" */
inline Node::Role GetRole(const std::string role_str) {
  if (role_str == "worker") return Node::WORKER;
  if (role_str == "server") return Node::SERVER;
  if (role_str == "scheduler") return Node::SCHEDULER;
  if (role_str == "joint") return Node::JOINT;
  CHECK(false) << "Unexpected role: " << role_str;
  return Node::SCHEDULER;
}

namespace ps_detail {
inline int GroupSizeFromEnv() {
  const int g = GetEnv("DMLC_GROUP_SIZE", 1);
  return g < 1 ? 1 : g;
}
/*! \brief every (postoffice, instance-level rank) this process runs for `role` */
inline std::vector<std::pair<Postoffice*, std::pair<Node::Role, int>>> Instances(Node::Role role,
                                                                                int rank,
                                                                                int group_size) {
  std::vector<std::pair<Postoffice*, std::pair<Node::Role, int>>> out;
  if (role == Node::SCHEDULER) {
    out.push_back({Postoffice::GetScheduler(), {Node::SCHEDULER, rank}});
    return out;
  }
  for (int i = 0; i < group_size; ++i) {
    const int r = (group_size == 1) ? rank : rank * group_size + i;
    if (role == Node::WORKER || role == Node::JOINT)
      out.push_back({Postoffice::GetWorker(i), {Node::WORKER, r}});
    if (role == Node::SERVER || role == Node::JOINT)
      out.push_back({Postoffice::GetServer(i), {Node::SERVER, r}});
  }
  return out;
}
}  // namespace ps_detail

/*!
 * \brief start the system; blocks until every node has registered (and, with
 *        do_barrier, until every node has reached this call).
 * \param customer_id id of the calling customer (0 for the first / only one)
 * \param role this process's role
 * \param rank preferred group rank, -1 to let the scheduler choose (required >= 0
 *        when DMLC_GROUP_SIZE > 1)
 */
inline void StartPS(int customer_id, Node::Role role, int rank, bool do_barrier,
                    const char* argv0 = nullptr) {
  const int group_size = ps_detail::GroupSizeFromEnv();
  Postoffice::Init(role);
  if (group_size > 1 && role != Node::SCHEDULER) CHECK_GE(rank, 0) << "rank required with groups";
  auto insts = ps_detail::Instances(role, rank, group_size);
  if (insts.size() == 1) {
    insts[0].first->Start(customer_id, insts[0].second.first, insts[0].second.second, do_barrier,
                          argv0);
    return;
  }
  // every instance registers with the scheduler concurrently
  std::vector<std::thread> threads;
  for (auto& it : insts) {
    threads.emplace_back([=] {
      it.first->Start(customer_id, it.second.first, it.second.second, do_barrier, argv0);
    });
  }
  for (auto& t : threads) t.join();
}

/*! \brief tear the system down; every node must call it before exiting */
inline void Finalize(int customer_id, Node::Role role, const bool do_barrier = true) {
  const int group_size = ps_detail::GroupSizeFromEnv();
  auto insts = ps_detail::Instances(role, 0, group_size);
  if (insts.size() == 1) {
    insts[0].first->Finalize(customer_id, do_barrier);
    return;
  }
  std::vector<std::thread> threads;
  for (auto& it : insts) {
    threads.emplace_back([=] { it.first->Finalize(customer_id, do_barrier); });
  }
  for (auto& t : threads) t.join();
}

/*! \brief `cb` runs right after Finalize() completes */
inline void RegisterExitCallback(const std::function<void()>& cb) {
  Postoffice::Get()->RegisterExitCallback(cb);
}

}  // namespace ps
#endif  // PS_PS_H_
