/**
 * \file multi_van.h
 * \brief MultiVan: one logical van striped over DMLC_NUM_PORTS independent TCP rails.
 *
 * Counterpart of the reference's test-oriented MultiVan (src/multi_van.h:59-284:
 * N inner ZMQ vans, rail picked by the values' src_device_id_, peer port by
 * dst_device_id_). Used to exercise multi-port / multi-device routing without
 * GPUs, and as a multi-rail CPU transport: with R rails a node has R listening
 * ports (Node::ports[0..R)), every rail connects to every port of every peer, and
 * R receive threads feed one queue. Control messages always use rail 0.
 */
#ifndef PS_VAN_MULTI_VAN_H_
#define PS_VAN_MULTI_VAN_H_
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "ps/internal/threadsafe_queue.h"
#include "van/tcp_van.h"

namespace ps {

class MultiVan : public Van {
 public:
  explicit MultiVan(Postoffice* postoffice) : Van(postoffice), postoffice_ptr_(postoffice) {}
  ~MultiVan() override { StopRails(); }

  std::string GetType() const override { return "multivan"; }

  void Start(int customer_id, bool standalone) override {
    if (rails_.empty()) {
      num_rails_ = std::max(1, GetEnv("DMLC_NUM_PORTS", GetEnv("DMLC_NUM_CPU_DEV", 1)));
      CHECK_LE(num_rails_, kMaxNodePorts);
      for (int i = 0; i < num_rails_; ++i) rails_.emplace_back(new Rail(postoffice_ptr_));
    }
    Van::Start(customer_id, standalone);
  }

  void Stop() override {
    Van::Stop();
    StopRails();
  }

  void SetNode(const Node& node) override {
    Van::SetNode(node);
    for (auto& r : rails_) r->Identify(node);
  }

  void RegisterRecvBuffer(Message& msg) override {
    for (auto& r : rails_) r->RegisterRecvBuffer(msg);
  }

 protected:
  /*! \brief a TcpVan used purely as a transport (no control plane of its own) */
  class Rail : public TcpVan {
   public:
    explicit Rail(Postoffice* po) : TcpVan(po) {}
    // a rail is driven by the MultiVan's own receive logic, never by Van::Receiving
    bool AllowLocalHandoff() const override { return false; }
    void Open() { InitTransport(); }
    void Identify(const Node& n) { Van::SetNode(n); }
    int BindPort(Node& n, int retry) { return Bind(n, retry); }
    void ConnectTo(const Node& n) { Connect(n); }
    int SendRaw(Message& m) { return SendMsg(m); }
    int RecvRaw(Message* m) { return RecvMsg(m); }
    void Wake() {
      Message bye;
      bye.meta.control.cmd = Control::TERMINATE;
      Loopback(bye);
    }
  };

  /*! \brief peer `id` reached through its port `port_idx`, as a distinct id for a rail */
  static int Mangle(int id, int port_idx) { return 1000000 + id + port_idx * 10000; }

  int Bind(Node& node, int max_retry) override {
    CHECK(!rails_.empty());
    // the scheduler listens on a single well-known port
    const int n = node.role == Node::SCHEDULER ? 1 : num_rails_;
    node.num_ports = n;
    for (int i = 0; i < n; ++i) {
      rails_[i]->Open();
      Node sub = node;
      sub.port = node.ports[i] ? node.ports[i] : node.port;
      const int got = rails_[i]->BindPort(sub, max_retry);
      if (got < 0) return -1;
      node.ports[i] = got;
      node.dev_types[i] = CPU;
      node.dev_ids[i] = i;
    }
    for (int i = n; i < num_rails_; ++i) rails_[i]->Open();  // send-only rails
    node.port = node.ports[0];
    for (auto& r : rails_) r->Identify(node);
    // one pump thread per listening rail
    running_ = true;
    for (int i = 0; i < n; ++i) {
      pumps_.emplace_back([this, i] {
        for (;;) {
          Message m;
          const int bytes = rails_[i]->RecvRaw(&m);
          if (!running_.load() || bytes < 0) return;
          inbox_.Push(std::make_pair(bytes, std::move(m)));
        }
      });
    }
    return node.port;
  }

  void Connect(const Node& node) override {
    if (node.id == my_node_.id) return;
    if (node.role == my_node_.role && node.role != Node::SCHEDULER) return;
    const int peer_ports = std::max(1, node.num_ports);
    {
      std::lock_guard<std::mutex> lk(mu_);
      ports_of_[node.id] = peer_ports;
    }
    for (auto& r : rails_) {
      for (int j = 0; j < peer_ports; ++j) {
        Node sub = node;
        sub.id = Mangle(node.id, j);
        sub.port = node.ports[j] ? node.ports[j] : node.port;
        sub.role = node.role;
        r->ConnectTo(sub);
      }
    }
  }

  int SendMsg(Message& msg) override {
    const int recver = msg.meta.recver;
    if (recver == my_node_.id) {  // loopback through the queue
      Message copy = msg;
      copy.meta.sender = my_node_.id;
      inbox_.Push(std::make_pair(1, std::move(copy)));
      return 1;
    }
    int rail = 0, port = 0;
    if (msg.meta.control.empty() && msg.data.size() >= 2) {
      // data: rail by where the values live, peer port by where they should land
      rail = std::max(0, msg.data[1].src_device_id_) % num_rails_;
      int peer_ports = 1;
      {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = ports_of_.find(recver);
        if (it != ports_of_.end()) peer_ports = it->second;
      }
      port = std::max(0, msg.data[1].dst_device_id_) % peer_ports;
    }
    Message routed = msg;
    routed.meta.recver = Mangle(recver, port);
    return rails_[rail]->SendRaw(routed);
  }

  int RecvMsg(Message* msg) override {
    std::pair<int, Message> item;
    inbox_.WaitAndPop(&item);
    *msg = std::move(item.second);
    msg->meta.recver = my_node_.id;
    return item.first;
  }

 private:
  void StopRails() {
    if (!running_.exchange(false) && pumps_.empty()) return;
    // unblock each pump with a loopback message on its own rail
    for (size_t i = 0; i < pumps_.size(); ++i) rails_[i]->Wake();
    for (auto& t : pumps_) t.join();
    pumps_.clear();
    rails_.clear();
  }

  Postoffice* postoffice_ptr_;
  int num_rails_ = 1;
  std::vector<std::unique_ptr<Rail>> rails_;
  std::vector<std::thread> pumps_;
  std::atomic<bool> running_{false};
  ThreadsafeQueue<std::pair<int, Message>> inbox_;
  std::mutex mu_;
  std::unordered_map<int, int> ports_of_;
};

}  // namespace ps
#endif  // PS_VAN_MULTI_VAN_H_
