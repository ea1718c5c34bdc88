/**
 * \file postoffice.cc
 * \brief Postoffice implementation (see postoffice.h).
 * Behavioural parity with reference src/postoffice.cc:20-304; differences:
 * GetCustomer waits on a condition variable instead of polling every 1 ms.
 */
#include "ps/internal/postoffice.h"
#include "core/sampler.h"
#include <chrono>
#include <thread>
#include "ps/base.h"
#include "ps/internal/message.h"

namespace ps {

Postoffice* Postoffice::po_scheduler_ = nullptr;
std::mutex Postoffice::init_mu_;
std::vector<Postoffice*> Postoffice::po_worker_group_;
std::vector<Postoffice*> Postoffice::po_server_group_;

Postoffice* Postoffice::Get() {
  if (po_scheduler_) return po_scheduler_;
  if (!po_server_group_.empty()) return po_server_group_[0];
  CHECK(!po_worker_group_.empty()) << "Please call ps::StartPS() first";
  return po_worker_group_[0];
}
Postoffice* Postoffice::GetServer(int index) {
  if (po_server_group_.empty()) {
    CHECK(po_scheduler_) << "Please call ps::StartPS() first";
    return po_scheduler_;
  }
  return po_server_group_.at(index);
}
Postoffice* Postoffice::GetScheduler() { return po_scheduler_; }
Postoffice* Postoffice::GetWorker(int index) {
  CHECK(!po_worker_group_.empty()) << "Please call ps::StartPS() first";
  return po_worker_group_.at(index);
}

void Postoffice::Init(Node::Role role) {
  std::lock_guard<std::mutex> lk(init_mu_);
  const int group_size = std::max(1, GetEnv("DMLC_GROUP_SIZE", 1));
  if (role == Node::SCHEDULER && !po_scheduler_) po_scheduler_ = new Postoffice(0);
  if ((role == Node::WORKER || role == Node::JOINT) && po_worker_group_.empty()) {
    for (int i = 0; i < group_size; ++i) po_worker_group_.push_back(new Postoffice(i));
  }
  if ((role == Node::SERVER || role == Node::JOINT) && po_server_group_.empty()) {
    for (int i = 0; i < group_size; ++i) po_server_group_.push_back(new Postoffice(i));
  }
}

void Postoffice::Reset() {
  std::lock_guard<std::mutex> lk(init_mu_);
  delete po_scheduler_;
  po_scheduler_ = nullptr;
  for (auto* p : po_worker_group_) delete p;
  for (auto* p : po_server_group_) delete p;
  po_worker_group_.clear();
  po_server_group_.clear();
}

Postoffice::Postoffice(int instance_idx) : instance_idx_(instance_idx) {
  env_ref_ = Environment::_GetSharedRef();
}

Postoffice::~Postoffice() { delete van_; }

void Postoffice::InitEnvironment() {
  group_size_ = std::max(1, GetEnv("DMLC_GROUP_SIZE", 1));
  std::string van_type = GetEnvStr("DMLC_ENABLE_RDMA", "zmq");
  if (GetEnv("DMLC_ENABLE_UCX", 0)) van_type = "ucx";
  if (const char* v = Environment::Get()->find("PS_VAN_TYPE")) van_type = v;
  if (!van_) van_ = Van::Create(van_type, this);
  num_workers_ = atoi(CHECK_NOTNULL(Environment::Get()->find("DMLC_NUM_WORKER")));
  num_servers_ = atoi(CHECK_NOTNULL(Environment::Get()->find("DMLC_NUM_SERVER")));
  verbose_ = GetEnv("PS_VERBOSE", 0);
}

void Postoffice::BuildGroupTable() {
  node_ids_.clear();
  const int W = kWorkerGroup, S = kServerGroup, H = kScheduler;
  for (int i = 0; i < num_workers_ * group_size_; ++i) {
    const int id = WorkerRankToID(i);
    for (int g : {id, W, W + S, W + H, W + S + H}) node_ids_[g].push_back(id);
  }
  for (int i = 0; i < num_servers_ * group_size_; ++i) {
    const int id = ServerRankToID(i);
    for (int g : {id, S, W + S, S + H, W + S + H}) node_ids_[g].push_back(id);
  }
  for (int g : {H, H + S + W, H + W, H + S}) node_ids_[g].push_back(kScheduler);
}

void Postoffice::Start(int customer_id, const Node::Role role, int rank, const bool do_barrier,
                       const char* argv0) {
  SampleProfiler::StartIfRequested();  // PS_SAMPLE_PROFILE=<file>: CPU sampling of this process
  CHECK_GE(rank, -1);
  preferred_rank_ = rank;
  {
    std::lock_guard<std::mutex> lk(start_mu_);
    if (init_stage_ == 0) {
      InitEnvironment();
      CHECK(role == Node::WORKER || role == Node::SERVER || role == Node::SCHEDULER)
          << "a Postoffice instance has exactly one role, got " << role;
      is_worker_ = role == Node::WORKER;
      is_server_ = role == Node::SERVER;
      is_scheduler_ = role == Node::SCHEDULER;
      dmlc::InitLogging(argv0 ? argv0 : "ps-lite");
      BuildGroupTable();
      init_stage_ = 1;
    }
  }
  van_->Start(customer_id, false);
  {
    std::lock_guard<std::mutex> lk(start_mu_);
    if (init_stage_ == 1) {
      start_time_ = time(nullptr);
      init_stage_ = 2;
    }
  }
  if (do_barrier) DoBarrier(customer_id, kWorkerGroup + kServerGroup + kScheduler, true);
}

void Postoffice::Finalize(const int customer_id, const bool do_barrier) {
  if (do_barrier) DoBarrier(customer_id, kWorkerGroup + kServerGroup + kScheduler, true);
  if (customer_id == 0) {
    num_workers_ = 0;
    num_servers_ = 0;
    van_->Stop();
    init_stage_ = 0;
    {
      std::lock_guard<std::mutex> lk(mu_);
      customers_.clear();
    }
    node_ids_.clear();
    {
      std::lock_guard<std::mutex> lk(barrier_mu_);
      barrier_done_.clear();
    }
    server_key_ranges_.clear();
    {
      std::lock_guard<std::mutex> lk(heartbeat_mu_);
      heartbeats_.clear();
    }
    if (exit_callback_) exit_callback_();
  }
}

void Postoffice::AddCustomer(Customer* customer) {
  const int app_id = CHECK_NOTNULL(customer)->app_id();
  const int customer_id = customer->customer_id();
  {
    std::lock_guard<std::mutex> lk(mu_);
    CHECK_EQ(customers_[app_id].count(customer_id), (size_t)0)
        << "customer_id " << customer_id << " already exists";
    customers_[app_id][customer_id] = customer;
  }
  customer_cv_.notify_all();
  {
    std::lock_guard<std::mutex> blk(barrier_mu_);
    barrier_done_[app_id].emplace(customer_id, false);
  }
  if (van_) van_->DeliverParked();
}

void Postoffice::RemoveCustomer(Customer* customer) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    const int app_id = CHECK_NOTNULL(customer)->app_id();
    auto it = customers_.find(app_id);
    if (it != customers_.end()) {
      it->second.erase(customer->customer_id());
      if (it->second.empty()) customers_.erase(it);
    }
  }
  // nobody can look the customer up any more; let deliveries that already did finish
  std::unique_lock<std::shared_mutex> drain(deliver_mu_);
}

bool Postoffice::DeliverOwned(int app_id, int customer_id, Message* msg) {
  std::shared_lock<std::shared_mutex> in_flight(deliver_mu_);
  Customer* obj = GetCustomer(app_id, customer_id, 0);
  if (!obj) return false;
  obj->Accept(std::move(*msg));
  return true;
}

bool Postoffice::Deliver(int app_id, int customer_id, const Message& msg) {
  std::shared_lock<std::shared_mutex> in_flight(deliver_mu_);
  Customer* obj = GetCustomer(app_id, customer_id, 0);
  if (!obj) return false;
  obj->Accept(msg);
  return true;
}

Customer* Postoffice::GetCustomer(int app_id, int customer_id, int timeout) const {
  std::unique_lock<std::mutex> lk(mu_);
  Customer* found = nullptr;
  auto probe = [&] {
    auto it = customers_.find(app_id);
    if (it == customers_.end()) return false;
    auto jt = it->second.find(customer_id);
    if (jt == it->second.end()) return false;
    found = jt->second;
    return true;
  };
  if (timeout <= 0) {
    probe();
  } else {
    customer_cv_.wait_for(lk, std::chrono::seconds(timeout), probe);
  }
  return found;
}

void Postoffice::DoBarrier(int customer_id, int node_group, bool instance_barrier) {
  const int members = static_cast<int>(GetNodeIDs(node_group).size());
  // nothing to wait for: a single instance, or a single group
  if (instance_barrier && members <= 1) return;
  if (!instance_barrier && members <= group_size_) return;
  const auto role = van_->my_node().role;
  if (role == Node::SCHEDULER) {
    CHECK(node_group & kScheduler);
  } else if (role == Node::WORKER) {
    CHECK(node_group & kWorkerGroup);
  } else if (role == Node::SERVER) {
    CHECK(node_group & kServerGroup);
  }
  std::unique_lock<std::mutex> ulk(barrier_mu_);
  barrier_done_[0][customer_id] = false;
  Message req;
  req.meta.recver = kScheduler;
  req.meta.request = true;
  req.meta.control.cmd = instance_barrier ? Control::INSTANCE_BARRIER : Control::BARRIER;
  req.meta.app_id = 0;
  req.meta.customer_id = customer_id;
  req.meta.control.barrier_group = node_group;
  req.meta.timestamp = van_->GetTimestamp();
  CHECK_GT(van_->Send(req), 0);
  // a barrier that never releases means some member never arrived (crashed, or still busy):
  // name the barrier every PS_WAIT_WARN_S seconds instead of hanging silently
  static const int warn_s = GetEnv("PS_WAIT_WARN_S", 60);
  auto released = [this, customer_id] { return barrier_done_[0][customer_id]; };
  if (warn_s <= 0) {
    barrier_cond_.wait(ulk, released);
    return;
  }
  while (!barrier_cond_.wait_for(ulk, std::chrono::seconds(warn_s), released)) {
    LOG(WARNING) << "node " << van_->my_node().id << " (" << role_str() << ") still inside "
                 << (instance_barrier ? "instance " : "") << "barrier of group " << node_group << " after "
                 << warn_s << " s: a member has not arrived";
  }
}

void Postoffice::Barrier(int customer_id, int node_group) {
  DoBarrier(customer_id, node_group, false);
}

const std::vector<Range>& Postoffice::GetServerKeyRanges() {
  std::lock_guard<std::mutex> lk(server_key_ranges_mu_);
  if (server_key_ranges_.empty()) {
    const uint64_t span = kMaxKey / static_cast<uint64_t>(std::max(1, num_servers_));
    for (int i = 0; i < num_servers_; ++i) {
      server_key_ranges_.push_back(Range(span * i, span * (i + 1)));
    }
  }
  return server_key_ranges_;
}

void Postoffice::Manage(const Message& recv) {
  CHECK(!recv.meta.control.empty());
  const auto& ctrl = recv.meta.control;
  const bool is_barrier = ctrl.cmd == Control::BARRIER || ctrl.cmd == Control::INSTANCE_BARRIER;
  if (is_barrier && !recv.meta.request) {
    {
      std::lock_guard<std::mutex> lk(barrier_mu_);
      // The scheduler echoes the customer id of whichever request completed the
      // count, which need not be ours, so every waiter of the app is released;
      // a waiter re-arms its flag before its next request.
      for (auto& kv : barrier_done_[recv.meta.app_id]) kv.second = true;
    }
    barrier_cond_.notify_all();
  }
}

std::vector<int> Postoffice::GetDeadNodes(int t) {
  std::vector<int> dead;
  if (!van_->IsReady() || t == 0) return dead;
  const time_t now = time(nullptr);
  const auto& nodes =
      is_scheduler_ ? GetNodeIDs(kWorkerGroup + kServerGroup) : GetNodeIDs(kScheduler);
  std::lock_guard<std::mutex> lk(heartbeat_mu_);
  for (int r : nodes) {
    auto it = heartbeats_.find(r);
    const bool silent = it == heartbeats_.end() || it->second + t < now;
    if (silent && start_time_ + t < now) dead.push_back(r);
  }
  return dead;
}

}  // namespace ps
