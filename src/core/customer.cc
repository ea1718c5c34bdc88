/**
 * \file customer.cc
 * \brief Customer inbox + ring-recycled request tracker (see customer.h).
 */
#include "ps/internal/customer.h"
#include <chrono>
#include <limits>
#include "ps/internal/postoffice.h"

namespace ps {

const int Node::kEmpty = std::numeric_limits<short>::max();
const int Meta::kEmpty = std::numeric_limits<short>::max();

namespace {
const size_t kInitialRing = 1024;
}

Customer::Customer(int app_id, int customer_id, const RecvHandle& recv_handle,
                   Postoffice* postoffice, bool start_now)
    : app_id_(app_id), customer_id_(customer_id), recv_handle_(recv_handle),
      postoffice_(postoffice) {
  ring_.resize(kInitialRing);
  direct_dispatch_ = GetEnv("PS_DIRECT_DISPATCH", 0) != 0;
  if (start_now) Start();
}

void Customer::Start() {
  if (started_) return;
  started_ = true;
  postoffice_->AddCustomer(this);
  if (!direct_dispatch_) {
    recv_thread_.reset(new std::thread(&Customer::Receiving, this));
  }
}

Customer::~Customer() {
  if (started_) postoffice_->RemoveCustomer(this);
  if (recv_thread_) {
    Message bye;
    bye.meta.control.cmd = Control::TERMINATE;
    inbox_.Push(bye);
    recv_thread_->join();
  }
}

Customer::Slot* Customer::Find(int ts) {
  Slot& s = ring_[static_cast<size_t>(ts) & (ring_.size() - 1)];
  return s.ts == ts ? &s : nullptr;
}

int Customer::NewRequest(int recver, int num_expected) {
  // requests fan out to one instance of every group in `recver`
  int groups = num_expected;
  if (groups < 0) {
    const int members = static_cast<int>(postoffice_->GetNodeIDs(recver).size());
    const bool is_group = recver < 8;
    groups = is_group ? std::max(1, members / postoffice_->group_size()) : members;
  }
  std::lock_guard<std::mutex> lk(tracker_mu_);
  // Meta::kEmpty (32767) means "no timestamp" in several places: never hand it out (the
  // reference does, and its resender then aborts on the 32768th request of a customer)
  if (next_ts_ == Meta::kEmpty) ++next_ts_;
  if (next_ts_ < 0) next_ts_ = 0;  // wrapped after 2^31 requests
  const int ts = next_ts_++;
  for (;;) {
    Slot& s = ring_[static_cast<size_t>(ts) & (ring_.size() - 1)];
    if (s.ts < 0 || s.received >= s.expected) {
      s.ts = ts;
      s.expected = groups;
      s.received = 0;
      break;
    }
    // the slot we would recycle is still in flight: double the ring and re-home
    std::vector<Slot> bigger(ring_.size() * 2);
    for (const Slot& old : ring_) {
      if (old.ts >= 0) bigger[static_cast<size_t>(old.ts) & (bigger.size() - 1)] = old;
    }
    ring_.swap(bigger);
  }
  return ts;
}

void Customer::WaitRequest(int timestamp) {
  std::unique_lock<std::mutex> lk(tracker_mu_);
  auto done = [this, timestamp] {
    Slot* s = Find(timestamp);
    // a recycled slot means the request completed long ago
    return s == nullptr || s->received >= s->expected;
  };
  // a request that never completes is the usual face of a transport bug or a dead peer: say
  // which one it is instead of hanging silently (PS_WAIT_WARN_S seconds, 0 = never)
  static const int warn_s = GetEnv("PS_WAIT_WARN_S", 60);
  if (warn_s <= 0) {
    tracker_cv_.wait(lk, done);
    return;
  }
  while (!tracker_cv_.wait_for(lk, std::chrono::seconds(warn_s), done)) {
    Slot* s = Find(timestamp);
    LOG(WARNING) << "app " << app_id_ << " customer " << customer_id_ << " on node "
                 << postoffice_->van()->my_node().id << ": request " << timestamp << " still waiting after "
                 << warn_s << " s (" << (s ? s->received : -1) << " of " << (s ? s->expected : -1)
                 << " responses)";
  }
}

int Customer::NumResponse(int timestamp) {
  std::lock_guard<std::mutex> lk(tracker_mu_);
  Slot* s = Find(timestamp);
  return s ? s->received : 0;
}

void Customer::AddResponse(int timestamp, int num) {
  {
    std::lock_guard<std::mutex> lk(tracker_mu_);
    Slot* s = Find(timestamp);
    if (s) s->received += num;
  }
  tracker_cv_.notify_all();
}

void Customer::Deliver(const Message& m) {
  recv_handle_(m);
  if (!m.meta.request) {
    {
      std::lock_guard<std::mutex> lk(tracker_mu_);
      Slot* s = Find(m.meta.timestamp);
      if (s) ++s->received;
    }
    tracker_cv_.notify_all();
  }
}

void Customer::Accept(const Message& recved) {
  if (direct_dispatch_) {
    Deliver(recved);
  } else {
    inbox_.Push(recved);
  }
}

void Customer::Accept(Message&& recved) {
  if (direct_dispatch_) {
    Deliver(recved);
  } else {
    inbox_.Push(std::move(recved));
  }
}

void Customer::Receiving() {
  // PS_COALESCE_LAUNCHES: handle everything that is already queued under one cork, so that
  // the copies the handlers issue (pull replies, acks gated on kernels) share launches / events
  const bool coalesce = GetEnv("PS_COALESCE_LAUNCHES", 0) != 0;
  for (;;) {
    Message m;
    inbox_.WaitAndPop(&m);
    if (!m.meta.control.empty() && m.meta.control.cmd == Control::TERMINATE) break;
    if (!coalesce) {
      Deliver(m);
      continue;
    }
    bool stop = false;
    {
      Van::CorkScope cork(postoffice_->van());
      Deliver(m);
      for (int n = 1; n < 32 && inbox_.TryPop(&m); ++n) {
        if (!m.meta.control.empty() && m.meta.control.cmd == Control::TERMINATE) {
          stop = true;
          break;
        }
        Deliver(m);
      }
    }
    if (stop) break;
  }
}

}  // namespace ps
