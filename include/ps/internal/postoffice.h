/**
 * \file postoffice.h
 * \brief Postoffice: the per-instance hub (role, id maps, customers, barriers,
 *        heartbeats, server key ranges) plus the static instance groups.
 *
 * API parity: reference include/ps/internal/postoffice.h:18-312. Node ids:
 * scheduler 1, server rank r -> 2r+8, worker rank r -> 2r+9; group ids are the
 * kScheduler/kServerGroup/kWorkerGroup bitmasks. One process may host a
 * scheduler *and* worker/server instances (used by single-process tests and by
 * the torchrun launcher where rank 0 also schedules) — the reference cannot.
 */
#ifndef PS_INTERNAL_POSTOFFICE_H_
#define PS_INTERNAL_POSTOFFICE_H_
#include <algorithm>
#include <condition_variable>
#include <ctime>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "ps/internal/customer.h"
#include "ps/internal/env.h"
#include "ps/internal/van.h"
#include "ps/range.h"

namespace ps {

class Postoffice {
 public:
  /*! \brief first live instance in the order scheduler, server[0], worker[0] */
  static Postoffice* Get();
  /*! \brief server instance `index` (the scheduler instance in a scheduler-only process) */
  static Postoffice* GetServer(int index = 0);
  static Postoffice* GetScheduler();
  static Postoffice* GetWorker(int index = 0);
  /*! \brief create the instance group(s) for `role`; idempotent per role */
  static void Init(Node::Role role);
  /*! \brief destroy every instance (after Finalize); lets a process StartPS again */
  static void Reset();

  Van* van() { return van_; }

  void Start(int customer_id, const Node::Role role, int rank, const bool do_barrier,
             const char* argv0);
  void Finalize(const int customer_id, const bool do_barrier = true);

  void AddCustomer(Customer* customer);
  void RemoveCustomer(Customer* customer);
  /*! \brief lookup; waits up to `timeout` seconds for the customer to be created */
  Customer* GetCustomer(int app_id, int customer_id, int timeout = 0) const;
  /*!
   * \brief hand `msg` to the customer if it exists (false otherwise). Unlike
   *        GetCustomer()->Accept() this cannot race with the customer's destruction:
   *        RemoveCustomer waits for deliveries in flight.
   */
  bool Deliver(int app_id, int customer_id, const Message& msg);
  /*! \brief same; `*msg` is moved from only when true is returned */
  bool DeliverOwned(int app_id, int customer_id, Message* msg);

  /*! \brief instance ids of a group id, or {id} for a single node id */
  const std::vector<int>& GetNodeIDs(int node_id) const {
    const auto it = node_ids_.find(node_id);
    CHECK(it != node_ids_.cend()) << "node " << node_id << " doesn't exist";
    return it->second;
  }
  /*! \brief uniform split of [0, kMaxKey) over the server *groups* */
  const std::vector<Range>& GetServerKeyRanges();

  using Callback = std::function<void()>;
  void RegisterExitCallback(const Callback& cb) { exit_callback_ = cb; }

  int GroupWorkerRankToInstanceID(int rank, int instance_idx) {
    return WorkerRankToID(rank * group_size_ + instance_idx);
  }
  int GroupServerRankToInstanceID(int rank, int instance_idx) {
    return ServerRankToID(rank * group_size_ + instance_idx);
  }
  int InstanceIDtoGroupRank(int id) { return IDtoRank(id) / group_size_; }
  static inline int WorkerRankToID(int rank) { return rank * 2 + 9; }
  static inline int ServerRankToID(int rank) { return rank * 2 + 8; }
  static inline int IDtoRank(int id) { return std::max((id - 8) / 2, 0); }

  int group_size() const { return group_size_; }
  int num_workers() const { return num_workers_; }
  int num_servers() const { return num_servers_; }
  int num_worker_instances() const { return num_workers_ * group_size_; }
  int num_server_instances() const { return num_servers_ * group_size_; }
  /*! \brief instance-level rank of this node in its role */
  int my_rank() const { return IDtoRank(van_->my_node().id); }
  int preferred_rank() const { return preferred_rank_; }
  int instance_idx() const { return instance_idx_; }
  int is_worker() const { return is_worker_; }
  int is_server() const { return is_server_; }
  int is_scheduler() const { return is_scheduler_; }
  std::string role_str() const {
    return is_scheduler_ ? "scheduler" : (is_server_ ? "server" : "worker");
  }
  int verbose() const { return verbose_; }
  bool is_recovery() const { return van_->my_node().is_recovery; }

  /*! \brief group-level barrier over `node_group` (any OR of the group ids) */
  void Barrier(int customer_id, int node_group);
  /*! \brief control-message sink called by the van (barrier releases) */
  void Manage(const Message& recv);
  void UpdateHeartbeat(int node_id, time_t t) {
    std::lock_guard<std::mutex> lk(heartbeat_mu_);
    heartbeats_[node_id] = t;
  }
  /*! \brief nodes with no heartbeat in the last `t` seconds */
  std::vector<int> GetDeadNodes(int t = 60);

 private:
  explicit Postoffice(int instance_idx);
  ~Postoffice();
  void InitEnvironment();
  void BuildGroupTable();
  void DoBarrier(int customer_id, int node_group, bool instance_barrier);

  static Postoffice* po_scheduler_;
  static std::mutex init_mu_;
  static std::vector<Postoffice*> po_worker_group_;
  static std::vector<Postoffice*> po_server_group_;

  Van* van_ = nullptr;
  mutable std::mutex mu_;
  std::shared_mutex deliver_mu_;  // shared: a delivery in flight; exclusive: a customer leaving
  mutable std::condition_variable customer_cv_;
  std::unordered_map<int, std::unordered_map<int, Customer*>> customers_;
  std::unordered_map<int, std::vector<int>> node_ids_;
  std::mutex server_key_ranges_mu_;
  std::vector<Range> server_key_ranges_;
  bool is_worker_ = false, is_server_ = false, is_scheduler_ = false;
  int num_servers_ = 0, num_workers_ = 0, group_size_ = 1;
  int preferred_rank_ = -1;
  int verbose_ = 0;
  // (app_id, customer_id) -> released?
  std::unordered_map<int, std::unordered_map<int, bool>> barrier_done_;
  std::mutex barrier_mu_;
  std::condition_variable barrier_cond_;
  std::mutex heartbeat_mu_;
  std::unordered_map<int, time_t> heartbeats_;
  std::mutex start_mu_;
  int init_stage_ = 0;
  int instance_idx_ = 0;
  Callback exit_callback_;
  std::shared_ptr<Environment> env_ref_;
  time_t start_time_ = 0;
  Postoffice(const Postoffice&) = delete;
  Postoffice& operator=(const Postoffice&) = delete;
};

/*! \brief verbosity-gated log (PS_VERBOSE) */
#define PS_VLOG(x) LOG_IF(INFO, (x) <= ::ps::Postoffice::Get()->verbose())

}  // namespace ps
#endif  // PS_INTERNAL_POSTOFFICE_H_
