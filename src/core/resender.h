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
#include <deque>
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
      if (!duplicated) {
        // a retransmission follows its original within a few timeouts: remembering the last
        // kSeenWindow messages is enough, and keeps a long job's memory bounded
        seen_order_.push_back(sig);
        if (seen_order_.size() > kSeenWindow) {
          seen_.erase(seen_order_.front());
          seen_order_.pop_front();
        }
      }
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
  /*!
   * \brief identity of a message on both ends of a link: a 64-bit mix of (app, customer, sender,
   *        receiver, full 32-bit timestamp, request bit). The reference packs truncated fields
   *        (8-bit node ids, src/resender.h:98-100); any fixed packing of these fields into 64
   *        bits has to drop something that eventually repeats — a long job wraps a 17-bit
   *        timestamp after 131072 requests and would see its own new messages as duplicates.
   */
  uint64_t Signature(const Message& msg) {
    CHECK_NE(msg.meta.timestamp, Meta::kEmpty) << msg.DebugString();
    const int s = msg.meta.sender == Node::kEmpty ? van_->my_node_.id : msg.meta.sender;
    uint64_t h = Mix(static_cast<uint32_t>(msg.meta.timestamp));
    h = Mix(h ^ (static_cast<uint64_t>(static_cast<uint32_t>(s)) << 32 | static_cast<uint32_t>(msg.meta.recver)));
    h = Mix(h ^ (static_cast<uint64_t>(static_cast<uint32_t>(msg.meta.app_id)) << 32 |
                 static_cast<uint32_t>(msg.meta.customer_id)));
    return Mix(h ^ (msg.meta.request ? 0x9e3779b97f4a7c15ull : 0));
  }
  static uint64_t Mix(uint64_t x) {  // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
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
          CHECK_LT(p.retries, max_retry_)
              << "no ACK after " << p.retries << " retransmissions over " << (now - p.sent_at) << " ms "
              << "(PS_RESEND_TIMEOUT=" << timeout_ms_ << ", PS_RESEND_MAX_RETRY=" << max_retry_
              << "): the peer is gone, or so loaded that the timeout is too short. " << p.msg.DebugString();
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
  static constexpr size_t kSeenWindow = 1u << 18;
  std::unordered_set<uint64_t> seen_;
  std::deque<uint64_t> seen_order_;  // insertion order of seen_, oldest first
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
