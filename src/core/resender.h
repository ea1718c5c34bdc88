/**
 * \file resender.h
 * \brief Optional at-least-once delivery (PS_RESEND=1).
 *
 * Every non-ACK message is remembered under a 64-bit signature until the peer
 * ACKs it; a monitor thread retransmits entries older than
 * timeout * (1 + retries) and gives up loudly after `max_retry`. The receive
 * side ACKs everything and filters duplicates.
 * Parity: reference src/resender.h:15-139. Fixes: the signature keeps 16 bits
 * of sender and receiver id (the reference truncates to 8, colliding past ~120
 * nodes), and the monitor wakes promptly on shutdown.
 */
#ifndef PS_CORE_RESENDER_H_
#define PS_CORE_RESENDER_H_
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include "ps/internal/van.h"

namespace ps {

class Resender {
 public:
  Resender(int timeout_ms, int max_retry, Van* van)
      : timeout_ms_(timeout_ms), max_retry_(max_retry), van_(van) {
    monitor_ = std::thread(&Resender::Monitor, this);
  }
  ~Resender() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      exit_ = true;
    }
    cv_.notify_all();
    monitor_.join();
  }

  /*! \brief remember an outgoing message until it is ACKed */
  void AddOutgoing(const Message& msg) {
    if (msg.meta.control.cmd == Control::ACK) return;
    // before the scheduler assigned our id nobody could address an ACK to us
    if (van_->my_node_.id == Node::kEmpty || msg.meta.recver == van_->my_node_.id) return;
    CHECK_NE(msg.meta.timestamp, Meta::kEmpty) << msg.DebugString();
    const uint64_t sig = Signature(msg);
    std::lock_guard<std::mutex> lk(mu_);
    if (pending_.count(sig)) return;  // a retransmission
    Pending& p = pending_[sig];
    p.msg = msg;
    p.sent_at = NowMs();
    p.retries = 0;
  }

  /*! \brief returns true if the message must not be processed (ACK or duplicate) */
  bool AddIncomming(const Message& msg) {
    if (msg.meta.control.cmd == Control::TERMINATE) return false;
    // registration traffic of a node without an id, and self-addressed messages
    if (msg.meta.sender == Node::kEmpty || msg.meta.sender == van_->my_node_.id) return false;
    if (msg.meta.control.cmd == Control::ACK) {
      std::lock_guard<std::mutex> lk(mu_);
      pending_.erase(msg.meta.control.msg_sig);
      return true;
    }
    const uint64_t sig = Signature(msg);
    bool duplicated;
    {
      std::lock_guard<std::mutex> lk(mu_);
      duplicated = !seen_.insert(sig).second;
    }
    Message ack;
    ack.meta.recver = msg.meta.sender;
    ack.meta.sender = msg.meta.recver;
    ack.meta.control.cmd = Control::ACK;
    ack.meta.control.msg_sig = sig;
    van_->Send(ack);
    if (duplicated) LOG(WARNING) << "Duplicated message: " << msg.DebugString();
    return duplicated;
  }

  /*! \brief wait (bounded) until every message we sent has been ACKed */
  void Flush(int deadline_ms) {
    const int64_t until = NowMs() + deadline_ms;
    while (NowMs() < until) {
      if (NumPending() == 0) return;
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
  }
  /*! \brief shutting down: stop treating unreachable peers as fatal */
  void SetLenient() {
    std::lock_guard<std::mutex> lk(mu_);
    lenient_ = true;
  }

  size_t NumPending() {
    std::lock_guard<std::mutex> lk(mu_);
    return pending_.size();
  }

 private:
  struct Pending {
    Message msg;
    int64_t sent_at = 0;
    int retries = 0;
  };
  static int64_t NowMs() {
    return std::chrono::duration_cast<std::chrono::milliseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  /*! \brief [app:14][sender:16][recver:16][timestamp:17][request:1] */
  uint64_t Signature(const Message& msg) {
    CHECK_NE(msg.meta.timestamp, Meta::kEmpty) << msg.DebugString();
    const uint64_t app = static_cast<uint64_t>(msg.meta.app_id) & 0x3fff;
    const int s = msg.meta.sender == Node::kEmpty ? van_->my_node_.id : msg.meta.sender;
    const uint64_t sender = static_cast<uint64_t>(s) & 0xffff;
    const uint64_t recver = static_cast<uint64_t>(msg.meta.recver) & 0xffff;
    const uint64_t ts = static_cast<uint64_t>(msg.meta.timestamp) & 0x1ffff;
    return (app << 50) | (sender << 34) | (recver << 18) | (ts << 1) |
           (msg.meta.request ? 1u : 0u);
  }
  void Monitor() {
    std::unique_lock<std::mutex> lk(mu_);
    while (!exit_) {
      cv_.wait_for(lk, std::chrono::milliseconds(timeout_ms_));
      if (exit_) break;
      std::vector<Message> resend;
      const int64_t now = NowMs();
      for (auto it = pending_.begin(); it != pending_.end();) {
        Pending& p = it->second;
        if (p.sent_at + static_cast<int64_t>(timeout_ms_) * (1 + p.retries) < now) {
          if (lenient_ && p.retries + 1 >= max_retry_) {
            it = pending_.erase(it);  // the peer has left; nobody will ever ACK this
            continue;
          }
          resend.push_back(p.msg);
          ++p.retries;
          LOG(WARNING) << van_->my_node_.ShortDebugString()
                       << ": Timeout to get the ACK message. Resend (retry=" << p.retries
                       << ") " << p.msg.DebugString();
          CHECK_LT(p.retries, max_retry_);
        }
        ++it;
      }
      lk.unlock();
      std::vector<uint64_t> unreachable;
      for (auto& m : resend) {
        // a retransmission that cannot even be handed to the transport means the peer
        // has left (it got the original, our copy of its ACK was lost): stop insisting
        if (van_->SendBestEffort(m) < 0) unreachable.push_back(Signature(m));
      }
      lk.lock();
      for (uint64_t sig : unreachable) pending_.erase(sig);
    }
  }

  std::thread monitor_;
  std::unordered_map<uint64_t, Pending> pending_;
  std::unordered_set<uint64_t> seen_;
  bool exit_ = false;
  bool lenient_ = false;
  std::mutex mu_;
  std::condition_variable cv_;
  int timeout_ms_;
  int max_retry_;
  Van* van_;
};

}  // namespace ps
#endif  // PS_CORE_RESENDER_H_
