/**
 * \file van.cc
 * \brief Transport-independent control plane (see van.h).
 *
 * Behavioural parity with the reference's src/van.cc (registration :112-290,
 * id adoption :292-332, heartbeat :334-349, instance barrier :351-379, group
 * barrier :382-426, data dispatch + profiling :428-458, Start :484-602, Stop
 * :604-632, receive loop :643-687). The implementation is organised differently:
 * the scheduler keeps a registration book (registrants_) with explicit
 * collect / order / assign / recover steps, the rank ordering uses a proper strict
 * weak order, the drop-injection PRNG is seeded once, startup waits with 1 ms
 * granularity instead of 100 ms, and profiling state is per van instance.
 */
#include "ps/internal/van.h"

#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <sstream>
#include <thread>

#include "core/network_utils.h"
#include "core/resender.h"
#include "core/wire.h"
#include "ps/internal/customer.h"
#include "ps/internal/postoffice.h"
#include "van/van_factory.h"

namespace ps {

Van* Van::Create(const std::string& type, Postoffice* postoffice) {
  return CreateVanByType(type, postoffice);
}

Van::Van(Postoffice* postoffice) : postoffice_(postoffice) {}

Van::~Van() {}

void Van::PackMeta(const Meta& meta, std::vector<char>* buf) { wire::PackMeta(meta, buf); }
bool Van::UnpackMeta(const char* buf, size_t size, Meta* meta) {
  return wire::UnpackMeta(buf, size, meta);
}

// ---------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------

void Van::Start(int customer_id, bool standalone) {
  {
    std::lock_guard<std::mutex> lk(start_mu_);
    if (init_stage_ == 0) {
      stopping_ = false;
      is_scheduler_ = postoffice_->is_scheduler();
      if (!standalone) {
        scheduler_.hostname = CHECK_NOTNULL(Environment::Get()->find("DMLC_PS_ROOT_URI"));
        scheduler_.port = atoi(CHECK_NOTNULL(Environment::Get()->find("DMLC_PS_ROOT_PORT")));
        scheduler_.num_ports = 1;
        scheduler_.ports[0] = scheduler_.port;
        scheduler_.dev_types[0] = CPU;
        scheduler_.dev_ids[0] = 0;
        scheduler_.role = Node::SCHEDULER;
        scheduler_.id = kScheduler;
      }
      if (is_scheduler_ && !standalone) {
        Node me = scheduler_;
        me.pid = static_cast<int>(getpid());
        SetNode(me);
      } else if (!standalone) {
        Node me = my_node_;
        me.role = postoffice_->is_worker() ? Node::WORKER : Node::SERVER;
        std::string ip = GetEnvStr("DMLC_NODE_HOST");
        if (ip.empty()) {
          std::string itf = GetEnvStr("DMLC_INTERFACE");
          if (!itf.empty()) {
            GetIP(itf, &ip);
          } else {
            GetAvailableInterfaceAndIP(&itf, &ip);
          }
          if (ip.empty()) ip = "127.0.0.1";  // loopback-only sandbox
        }
        const int num_ports = std::max(1, GetEnv("DMLC_NUM_PORTS", GetEnv("DMLC_NUM_CPU_DEV", 1)));
        CHECK_LE(num_ports, kMaxNodePorts);
        std::array<int, 32> ports;
        ports.fill(0);
        const int got = GetAvailablePort(num_ports, &ports);
        CHECK_EQ(got, num_ports) << "failed to get " << num_ports << " ports";
        if (const char* p = Environment::Get()->find("DMLC_PORT")) ports[0] = atoi(p);
        me.hostname = ip;
        me.num_ports = num_ports;
        for (int i = 0; i < num_ports; ++i) me.ports[i] = ports[i];
        me.port = ports[0];
        me.id = Node::kEmpty;  // assigned by the scheduler
        me.customer_id = customer_id;
        me.pid = static_cast<int>(getpid());
        SetNode(me);
      }
      my_node_.port = Bind(my_node_, is_scheduler_ ? 0 : 40);
      CHECK_NE(my_node_.port, -1) << "bind failed";
      my_node_.ports[0] = my_node_.port;
      LOG_IF(INFO, postoffice_->verbose() >= 1) << "Bind to " << my_node_.DebugString();

      if (!standalone) Connect(scheduler_);

      if (const char* d = Environment::Get()->find("PS_DROP_MSG")) {
        drop_rate_ = atoi(d);
        drop_seed_ = static_cast<unsigned>(time(nullptr)) ^ static_cast<unsigned>(getpid());
      }
      heartbeat_timeout_ = GetEnv("PS_HEARTBEAT_TIMEOUT", 0);

      if (GetEnv("ENABLE_PROFILING", 0) && !is_scheduler_ && !standalone) {
        auto us = std::chrono::duration_cast<std::chrono::microseconds>(
                      std::chrono::system_clock::now().time_since_epoch()).count();
        const char* role = postoffice_->is_worker() ? "worker" : "server";
        std::string path;
        if (const char* pp = Environment::Get()->find("PROFILE_PATH")) {
          path = std::string(pp) + "_van_" + role;
        } else {
          path = std::string("pslite_profile_van_") + role + "_" + std::to_string(us);
        }
        if (postoffice_->instance_idx() > 0) path += "_" + std::to_string(postoffice_->instance_idx());
        profile_out_.open(path, std::ios::out | std::ios::trunc);
        profiling_ = profile_out_.is_open();
      }

      // The resender must exist before the first message can arrive: a request that is
      // processed without being recorded (and ACKed) would be processed *again* when its
      // sender retransmits it (the reference creates it after registration, src/van.cc:587).
      if (GetEnv("PS_RESEND", 0) != 0 && !resender_) {
        resender_ = new Resender(GetEnv("PS_RESEND_TIMEOUT", 1000), GetEnv("PS_RESEND_MAX_RETRY", 10), this);
      }
      receiver_thread_.reset(new std::thread(&Van::Receiving, this));
      if (standalone) ready_ = true;
      init_stage_ = 1;
    }
  }
  if (standalone) return;

  if (!is_scheduler_) {
    // announce this (node, customer) to the scheduler
    Message hello;
    Node me = my_node_;
    me.aux_id = postoffice_->preferred_rank();
    me.customer_id = customer_id;
    hello.meta.recver = kScheduler;
    hello.meta.control.cmd = Control::ADD_NODE;
    hello.meta.control.node.push_back(me);
    hello.meta.timestamp = timestamp_++;
    Send(hello);
  }
  while (!ready_.load()) std::this_thread::sleep_for(std::chrono::milliseconds(1));

  {
    std::lock_guard<std::mutex> lk(start_mu_);
    if (init_stage_ == 1) {
      if (!is_scheduler_ && GetEnv("PS_HEARTBEAT_INTERVAL", 0) > 0) {
        heartbeat_thread_.reset(new std::thread(&Van::HeartbeatLoop, this));
      }
      init_stage_ = 2;
    }
  }
}

void Van::Stop() {
  if (resender_) {
    // give in-flight messages a chance to be ACKed, then stop insisting
    resender_->SetLenient();
    resender_->Flush(GetEnv("PS_RESEND_TIMEOUT", 1000) * 4);
  }
  stopping_ = true;
  // wake the receive loop with a message to ourselves
  Message bye;
  bye.meta.control.cmd = Control::TERMINATE;
  bye.meta.recver = my_node_.id;
  bye.meta.customer_id = 0;
  const int ret = SendMsg(bye);
  CHECK_NE(ret, -1);
  if (receiver_thread_) receiver_thread_->join();
  receiver_thread_.reset();
  if (heartbeat_thread_) heartbeat_thread_->join();
  heartbeat_thread_.reset();
  delete resender_;
  resender_ = nullptr;
  init_stage_ = 0;
  ready_ = false;
  connected_.clear();
  alias_of_.clear();
  registrants_.clear();
  table_final_ = false;
  num_servers_ = num_workers_ = 0;
  send_bytes_ = 0;
  recv_bytes_ = 0;
  timestamp_ = 0;
  my_node_.id = Meta::kEmpty;
  instance_barrier_count_.clear();
  group_barrier_requests_.clear();
  if (profiling_) {
    std::lock_guard<std::mutex> lk(profile_mu_);
    profile_out_.flush();
    profile_out_.close();
    profiling_ = false;
  }
}

int Van::SendBestEffort(Message& msg) {
  const int n = SendMsg(msg);
  if (n > 0) send_bytes_ += static_cast<size_t>(n);
  return n;
}

int Van::Send(Message& msg) {
  // remember the message BEFORE it leaves: over a shared-memory ring the ACK can be back before
  // SendMsg returns, and an ACK for a message nobody remembers is dropped (the message would be
  // retransmitted a timeout later for nothing; the reference registers it after the send)
  if (resender_) resender_->AddOutgoing(msg);
  const int n = SendMsg(msg);
  if (n == -1 && (stopping_.load() || msg.meta.control.cmd == Control::ACK)) {
    // the peer already left (shutdown races with late ACKs / retransmissions)
    LOG(WARNING) << GetType() << " could not deliver to departed node " << msg.meta.recver;
    return -1;
  }
  CHECK_NE(n, -1) << GetType() << " sent -1 bytes";
  send_bytes_ += static_cast<size_t>(n);
  LOG_IF(INFO, postoffice_->verbose() >= 2)
      << GetType() << " " << my_node_.id << "\tsent: " << msg.DebugString();
  return n;
}

// ---------------------------------------------------------------------------
// receive loop
// ---------------------------------------------------------------------------

void Van::Receiving() {
  for (;;) {
    Message msg;
    const int n = RecvMsg(&msg);
    CHECK_NE(n, -1) << GetType() << " receive failed";
    // fault injection models a lossy *network*: loopback (self-addressed) traffic is exempt
    if (ready_.load() && drop_rate_ > 0 && msg.meta.control.cmd != Control::TERMINATE &&
        msg.meta.sender != my_node_.id) {
      if (static_cast<int>(rand_r(&drop_seed_) % 100) < drop_rate_) {
        LOG(WARNING) << "Drop message " << msg.DebugString();
        continue;
      }
    }
    recv_bytes_ += static_cast<size_t>(n);
    LOG_IF(INFO, postoffice_->verbose() >= 2)
        << GetType() << " " << my_node_.id << "\treceived: " << msg.DebugString();
    if (resender_ && resender_->AddIncomming(msg)) continue;

    if (msg.meta.control.empty()) {
      ProcessDataMsg(&msg);
      continue;
    }
    switch (msg.meta.control.cmd) {
      case Control::TERMINATE:
        OnTerminate();
        return;
      case Control::ADD_NODE:
        OnAddNode(&msg);
        break;
      case Control::BARRIER:
        OnBarrier(&msg);
        break;
      case Control::INSTANCE_BARRIER:
        OnInstanceBarrier(&msg);
        break;
      case Control::HEARTBEAT:
        OnHeartbeat(&msg);
        break;
      default:
        if (!HandleTransportControl(&msg)) {
          LOG(WARNING) << "Drop unknown typed message " << msg.DebugString();
        }
    }
  }
}

void Van::OnTerminate() {
  LOG_IF(INFO, postoffice_->verbose() >= 1) << my_node_.ShortDebugString() << " is stopped";
  ready_ = false;
}

void Van::ProcessDataMsg(Message* msg) {
  CHECK_NE(msg->meta.sender, Meta::kEmpty);
  CHECK_NE(msg->meta.recver, Meta::kEmpty);
  CHECK_NE(msg->meta.app_id, Meta::kEmpty);
  const int app_id = msg->meta.app_id;
  // servers run one customer per app; workers may run several
  const int customer_id = postoffice_->is_worker() ? msg->meta.customer_id : app_id;
  // what the profiling line needs, read before the message is handed over
  const bool log_it = profiling_ && !msg->data.empty() && msg->data[0].size() >= 2 && !msg->data[0].on_gpu();
  int key16 = 0;
  const bool is_push = msg->meta.push;
  if (log_it) {
    const unsigned char* k = reinterpret_cast<const unsigned char*>(msg->data[0].data());
    key16 = k[0] + 256 * k[1];
  }
  // Fast path: nothing is parked (only receive threads park, so the count cannot grow behind
  // this thread's back) and the customer exists.
  if (parked_n_.load(std::memory_order_acquire) != 0 || !postoffice_->DeliverOwned(app_id, customer_id, msg)) {
    // The application has not created this customer yet (e.g. it is still inside the
    // start-up barrier), or older messages are still parked. Park the message instead of
    // blocking the receive thread — the reference waits here for up to 5 s (src/van.cc:435),
    // during which no barrier release, ACK or heartbeat can be processed. The look-up is
    // repeated under parked_mu_: Postoffice::AddCustomer inserts the customer and THEN drains
    // under the same lock, so either this thread sees the customer now or the drain sees the
    // parked message. While anything is parked, new messages queue behind it (a sender's stream
    // keeps its order) and the queue is drained in order right here.
    std::lock_guard<std::mutex> lk(parked_mu_);
    if (parked_.empty() && postoffice_->DeliverOwned(app_id, customer_id, msg)) {
      // delivered after all
    } else {
      parked_.push_back(*msg);
      DrainParkedLocked();
      return;
    }
  }
  if (log_it) {
    auto us = std::chrono::duration_cast<std::chrono::microseconds>(
                  std::chrono::system_clock::now().time_since_epoch()).count();
    std::lock_guard<std::mutex> lk(profile_mu_);
    profile_out_ << key16 << "\t" << (postoffice_->is_worker() ? "worker" : "server")
                 << "_van_recv_" << (is_push ? "push" : "pull") << "\t" << us << "\n";
  }
}

bool Van::AcceptHandoff(Message* msg) {
  if (!ready_.load() || stopping_.load() || resender_ != nullptr || drop_rate_ > 0 || profiling_) {
    return false;
  }
  if (msg->meta.app_id == Meta::kEmpty) return false;
  const int customer_id = postoffice_->is_worker() ? msg->meta.customer_id : msg->meta.app_id;
  const size_t bytes = static_cast<size_t>(msg->meta.data_size > 0 ? msg->meta.data_size : 0) + 64;
  if (postoffice_->verbose() >= 2) {
    LOG(INFO) << GetType() << " " << my_node_.id << "\treceived (in-process): " << msg->DebugString();
  }
  if (!postoffice_->DeliverOwned(msg->meta.app_id, customer_id, msg)) return false;
  recv_bytes_ += bytes;
  return true;
}

void Van::DeliverParked() {
  std::lock_guard<std::mutex> lk(parked_mu_);
  DrainParkedLocked();
}

void Van::DrainParkedLocked() {
  // in arrival order; a message whose customer is still missing stays, and so does everything
  // behind it for the same customer (Deliver fails for those as well)
  std::vector<Message> keep;
  for (Message& m : parked_) {
    const int customer_id = postoffice_->is_worker() ? m.meta.customer_id : m.meta.app_id;
    if (!postoffice_->Deliver(m.meta.app_id, customer_id, m)) keep.push_back(std::move(m));
  }
  parked_.swap(keep);
  parked_n_.store(static_cast<int>(parked_.size()), std::memory_order_release);
}

// ---------------------------------------------------------------------------
// registration
// ---------------------------------------------------------------------------

void Van::OnAddNode(Message* msg) {
  if (is_scheduler_) {
    SchedulerCollect(msg);
    return;
  }
  auto& nodes = msg->meta.control.node;
  AdoptIdentity(nodes);
  {
    std::lock_guard<std::mutex> lk(cluster_mu_);
    for (const Node& n : nodes) {
      if (n.role != Node::SERVER && n.role != Node::WORKER) continue;
      bool known = false;
      for (Node& c : cluster_) {
        if (c.id == n.id) {
          c = n;  // a recovered node replaces its predecessor
          known = true;
        }
      }
      if (!known) cluster_.push_back(n);
    }
  }
  for (const Node& n : nodes) {
    const std::string addr = n.Address();
    if (!connected_.count(addr)) {
      Connect(n);
      connected_[addr] = n.id;
    }
    if (!n.is_recovery && n.role == Node::SERVER) ++num_servers_;
    if (!n.is_recovery && n.role == Node::WORKER) ++num_workers_;
  }
  LOG_IF(INFO, postoffice_->verbose() >= 1)
      << my_node_.ShortDebugString() << " is connected to others";
  ready_ = true;
}

void Van::AdoptIdentity(const std::vector<Node>& nodes) {
  for (const Node& n : nodes) {
    if (n.hostname == my_node_.hostname && n.port == my_node_.port) {
      if (Environment::Get()->find("DMLC_RANK") == nullptr || my_node_.id == Meta::kEmpty) {
        Node me = n;
        // fields only the local process knows survive the round trip
        me.pid = my_node_.pid;
        SetNode(me);
      }
    }
  }
}

void Van::SchedulerCollect(Message* msg) {
  auto& ctrl = msg->meta.control;
  CHECK_EQ(ctrl.node.size(), (size_t)1);
  const size_t expected =
      postoffice_->num_server_instances() + postoffice_->num_worker_instances();
  if (!table_final_) {
    registrants_.push_back(ctrl.node[0]);
    if (registrants_.size() == expected) {
      SchedulerAssignAndBroadcast();
    } else {
      LOG_IF(INFO, postoffice_->verbose() >= 1)
          << "AddNode (" << registrants_.size() << "/" << expected
          << "): " << registrants_.back().DebugString();
    }
  } else {
    SchedulerHandleRecovery(ctrl.node[0]);
  }
}

void Van::OrderRegistrants(std::vector<Node>* nodes) {
  const bool mixed = GetEnv("BYTEPS_ENABLE_MIXED_MODE", 0) != 0;
  const std::string ordered_hosts = GetEnvStr("BYTEPS_ORDERED_HOSTS");
  CHECK(!(mixed && !ordered_hosts.empty()))
      << "BYTEPS_ENABLE_MIXED_MODE and BYTEPS_ORDERED_HOSTS should not coexist";
  auto by_addr = [](const Node& a, const Node& b) {
    if (a.hostname != b.hostname) return a.hostname < b.hostname;
    return a.port < b.port;
  };
  if (mixed) {
    // places that run only a server sort before places that co-locate worker+server. The
    // reference counts nodes per IP (one machine = one place, src/van.cc:126-140); on one box all
    // nodes share the IP, so a place is a PROCESS here: a joint process registers its worker
    // and its server with the same pid.
    auto place = [](const Node& n) { return n.hostname + "#" + std::to_string(n.pid); };
    std::unordered_map<std::string, int> per_place;
    for (const Node& n : *nodes) {
      ++per_place[place(n)];
      CHECK_LE(per_place[place(n)], 2) << place(n);
    }
    std::stable_sort(nodes->begin(), nodes->end(), [&](const Node& a, const Node& b) {
      const int ca = per_place[place(a)], cb = per_place[place(b)];
      if (ca != cb) return ca < cb;
      return by_addr(a, b);
    });
    for (const Node& n : *nodes) {
      if (per_place[place(n)] == 1) CHECK_EQ(n.role, Node::SERVER) << n.DebugString();
    }
  } else if (!ordered_hosts.empty()) {
    // rank follows the position of the node's IP in the comma-separated list
    std::unordered_map<std::string, size_t> pos;
    std::stringstream ss(ordered_hosts);
    std::string item;
    size_t idx = 0;
    while (std::getline(ss, item, ',')) {
      const std::string ip = item.substr(0, item.find(':'));
      CHECK(!pos.count(ip)) << "Duplicate IP found in BYTEPS_ORDERED_HOSTS: " << ip;
      pos[ip] = idx++;
    }
    std::stable_sort(nodes->begin(), nodes->end(), [&](const Node& a, const Node& b) {
      return pos[a.hostname] < pos[b.hostname];
    });
  } else {
    std::sort(nodes->begin(), nodes->end(), by_addr);
  }
}

void Van::SchedulerAssignAndBroadcast() {
  OrderRegistrants(&registrants_);

  bool with_preferred = false;
  for (const Node& n : registrants_) with_preferred |= (n.aux_id != -1);
  if (with_preferred) {
    // preferred ranks must form exactly 0..N-1 within each role
    std::unordered_set<int> s_ranks, w_ranks;
    for (const Node& n : registrants_) {
      auto& set = n.role == Node::SERVER ? s_ranks : w_ranks;
      CHECK(n.role == Node::SERVER || n.role == Node::WORKER) << n.DebugString();
      CHECK(set.insert(n.aux_id).second) << "rank must be unique: " << n.DebugString();
    }
    const int ns = postoffice_->num_server_instances();
    const int nw = postoffice_->num_worker_instances();
    CHECK_EQ(s_ranks.size(), (size_t)ns);
    CHECK_EQ(w_ranks.size(), (size_t)nw);
    for (int i = 0; i < ns; ++i) CHECK(s_ranks.count(i)) << "missing server rank " << i;
    for (int i = 0; i < nw; ++i) CHECK(w_ranks.count(i)) << "missing worker rank " << i;
  }

  const time_t now = time(nullptr);
  for (Node& n : registrants_) {
    const bool is_server = n.role == Node::SERVER;
    const int rank = with_preferred ? n.aux_id : (is_server ? num_servers_ : num_workers_);
    const int id = is_server ? Postoffice::ServerRankToID(rank) : Postoffice::WorkerRankToID(rank);
    const std::string addr = n.Address();
    auto it = connected_.find(addr);
    if (it == connected_.end()) {
      CHECK_EQ(n.id, Node::kEmpty);
      n.id = id;
      Connect(n);
      postoffice_->UpdateHeartbeat(n.id, now);
      connected_[addr] = id;
      LOG_IF(INFO, postoffice_->verbose() >= 1) << "assign id=" << id << " to " << n.DebugString();
    } else {
      // second customer of an already-known endpoint: alias onto the first id
      alias_of_[id] = it->second;
      n.id = it->second;
    }
    if (is_server) ++num_servers_; else ++num_workers_;
  }

  Message table;
  table.meta.control.cmd = Control::ADD_NODE;
  table.meta.control.node = registrants_;
  table.meta.control.node.push_back(my_node_);
  for (int r : postoffice_->GetNodeIDs(kWorkerGroup + kServerGroup)) {
    if (alias_of_.count(r)) continue;
    table.meta.recver = r;
    table.meta.timestamp = timestamp_++;
    Send(table);
  }
  LOG_IF(INFO, postoffice_->verbose() >= 1) << "The scheduler is connected to " << num_workers_
                                            << " workers and " << num_servers_ << " servers";
  table_final_ = true;
  ready_ = true;
}

void Van::SchedulerHandleRecovery(const Node& reborn_in) {
  CHECK(ready_.load());
  auto dead = postoffice_->GetDeadNodes(heartbeat_timeout_);
  std::unordered_set<int> dead_set(dead.begin(), dead.end());
  Node reborn = reborn_in;
  bool matched = false;
  for (Node& slot : registrants_) {
    if (dead_set.count(slot.id) && slot.role == reborn.role) {
      reborn.id = slot.id;
      reborn.is_recovery = true;
      LOG_IF(INFO, postoffice_->verbose() >= 1)
          << "replace dead node " << slot.DebugString() << " by node " << reborn.DebugString();
      slot = reborn;
      matched = true;
      break;
    }
  }
  if (!matched) {
    LOG(WARNING) << "late ADD_NODE with no dead node of that role to replace: "
                 << reborn.DebugString();
    return;
  }
  Connect(reborn);
  connected_[reborn.Address()] = reborn.id;
  postoffice_->UpdateHeartbeat(reborn.id, time(nullptr));
  for (int r : postoffice_->GetNodeIDs(kWorkerGroup + kServerGroup)) {
    if (r != reborn.id && dead_set.count(r)) continue;  // never write to the dead
    Message back;
    back.meta.control.cmd = Control::ADD_NODE;
    if (r == reborn.id) {
      back.meta.control.node = registrants_;  // newcomer learns everybody
      back.meta.control.node.push_back(my_node_);
    } else {
      back.meta.control.node.push_back(reborn);  // survivors learn the newcomer
    }
    back.meta.recver = r;
    back.meta.timestamp = timestamp_++;
    Send(back);
  }
}

// ---------------------------------------------------------------------------
// heartbeats and barriers
// ---------------------------------------------------------------------------

void Van::OnHeartbeat(Message* msg) {
  const time_t now = time(nullptr);
  for (const Node& n : msg->meta.control.node) {
    postoffice_->UpdateHeartbeat(n.id, now);
    if (is_scheduler_) {
      Message echo;
      echo.meta.recver = n.id;
      echo.meta.control.cmd = Control::HEARTBEAT;
      echo.meta.control.node.push_back(my_node_);
      echo.meta.timestamp = timestamp_++;
      Send(echo);
    }
  }
}

void Van::HeartbeatLoop() {
  const int interval = GetEnv("PS_HEARTBEAT_INTERVAL", 0);
  int64_t slept_ms = 0;
  while (interval > 0 && ready_.load() && !stopping_.load()) {
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    slept_ms += 20;
    if (slept_ms < static_cast<int64_t>(interval) * 1000) continue;
    slept_ms = 0;
    if (!ready_.load() || stopping_.load()) break;
    Message beat;
    beat.meta.recver = kScheduler;
    beat.meta.control.cmd = Control::HEARTBEAT;
    beat.meta.control.node.push_back(my_node_);
    beat.meta.timestamp = timestamp_++;
    Send(beat);
  }
}

void Van::OnInstanceBarrier(Message* msg) {
  if (!msg->meta.request) {
    postoffice_->Manage(*msg);
    return;
  }
  // every *instance* of the group must check in; everyone is released
  if (instance_barrier_count_.empty()) instance_barrier_count_.resize(8, 0);
  const int group = msg->meta.control.barrier_group;
  CHECK_LT(group, 8);
  const auto& members = postoffice_->GetNodeIDs(group);
  if (++instance_barrier_count_[group] < static_cast<int>(members.size())) return;
  instance_barrier_count_[group] = 0;
  Message release;
  release.meta.request = false;
  release.meta.app_id = msg->meta.app_id;
  release.meta.customer_id = msg->meta.customer_id;
  release.meta.control.cmd = Control::INSTANCE_BARRIER;
  for (int r : members) {
    if (alias_of_.count(r)) continue;
    release.meta.recver = r;
    release.meta.timestamp = timestamp_++;
    CHECK_GT(Send(release), 0);
  }
}

void Van::OnBarrier(Message* msg) {
  if (!msg->meta.request) {
    postoffice_->Manage(*msg);
    return;
  }
  // one request per *group* (one instance speaks for its group); only the
  // requesters are released
  const int group = msg->meta.control.barrier_group;
  auto& waiting = group_barrier_requests_[group];
  waiting.push_back(msg->meta.sender);
  const int gs = postoffice_->group_size();
  const int members = static_cast<int>(postoffice_->GetNodeIDs(group).size());
  size_t expected;
  if (group == kScheduler) {
    expected = 1;
  } else if (group & kScheduler) {
    expected = static_cast<size_t>((members - 1) / gs + 1);
  } else {
    expected = static_cast<size_t>(members / gs);
  }
  if (waiting.size() < expected) return;
  Message release;
  release.meta.request = false;
  release.meta.app_id = msg->meta.app_id;
  release.meta.customer_id = msg->meta.customer_id;
  release.meta.control.cmd = Control::BARRIER;
  for (int r : waiting) {
    if (alias_of_.count(r)) continue;
    release.meta.recver = r;
    release.meta.timestamp = timestamp_++;
    CHECK_GT(Send(release), 0);
  }
  waiting.clear();
}

}  // namespace ps
