/**
 * \file customer.cc
 * \brief Customer inbox + ring-recycled request tracker (see customer.h).
 */
#include "ps/internal/customer.h"
#include <chrono>
#include <limits>
#include "ps/internal/postoffice.h"
#include "core/event_trace.h"

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
  if (postoffice_->verbose() >= 1) {
    LOG(INFO) << "customer " << customer_id_ << " of app " << app_id_ << ": Wait() went to sleep "
              << wait_sleeps_.load() << " times";
  }
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
  std::lock_guard<SpinMutex> lk(tracker_mu_);
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
      s.waiters = 0;  // a waiter of the recycled request finds its slot gone and never decrements
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
  EventTrace::Mark("wait", timestamp);
  std::unique_lock<SpinMutex> lk(tracker_mu_);
  auto done = [this, timestamp] {
    Slot* s = Find(timestamp);
    // a recycled slot means the request completed long ago
    return s == nullptr || s->received >= s->expected;
  };
  if (done()) return;
  // responses of a busy connection arrive microseconds apart: look a few more times before
  // paying for a sleep and a wake-up (PS_WAIT_SPIN_US)
  static const int spin_us = GetEnv("PS_WAIT_SPIN_US", 100);
  if (spin_us > 0) {
    // poll the completion counter, not the tracker: the customer thread needs tracker_mu_ for
    // every response and must not fight this thread for it
    // A short hot phase, then sched_yield between polls: the thread that completes the request
    // may have been put on THIS cpu by the scheduler (it wakes us with a futex, and wake-affine
    // placement then lets the woken spinner run in front of its own waker) — a waiter that only
    // executes `pause` would keep it off the cpu for the whole window.
    static const int hot_us = GetEnv("PS_WAIT_HOT_US", 2);
    const auto t0 = std::chrono::steady_clock::now();
    const auto hot_until = t0 + std::chrono::microseconds(hot_us);
    const auto until = t0 + std::chrono::microseconds(spin_us);
    uint64_t seen = completions_.load(std::memory_order_acquire);
    lk.unlock();
    bool hot = true;
    for (;;) {
      if (hot) {
        for (int i = 0; i < 16; ++i) ThreadsafeQueue<Message>::CpuRelax();
      } else {
        std::this_thread::yield();
      }
      const uint64_t now = completions_.load(std::memory_order_acquire);
      if (now != seen) {
        seen = now;
        lk.lock();
        if (done()) return;
        lk.unlock();
      }
      const auto t = std::chrono::steady_clock::now();
      if (t >= until) break;
      if (hot && t >= hot_until) hot = false;
    }
    lk.lock();
    if (done()) return;
  }
  // responses wake only the threads that wait for *their* request (a round of the benchmark
  // has 80 requests in flight: waking the application thread for each costs two system calls)
  if (Slot* s = Find(timestamp)) ++s->waiters;
  wait_sleeps_.fetch_add(1, std::memory_order_relaxed);
  // a request that never completes is the usual face of a transport bug or a dead peer: say
  // which one it is instead of hanging silently (PS_WAIT_WARN_S seconds, 0 = never)
  static const int warn_s = GetEnv("PS_WAIT_WARN_S", 60);
  if (warn_s <= 0) {
    tracker_cv_.wait(lk, done);
  } else {
    while (!tracker_cv_.wait_for(lk, std::chrono::seconds(warn_s), done)) {
      Slot* s = Find(timestamp);
      LOG(WARNING) << "app " << app_id_ << " customer " << customer_id_ << " on node "
                   << postoffice_->van()->my_node().id << ": request " << timestamp << " still waiting after "
                   << warn_s << " s (" << (s ? s->received : -1) << " of " << (s ? s->expected : -1)
                   << " responses)";
    }
  }
  if (Slot* s = Find(timestamp)) --s->waiters;
}

int Customer::NumResponse(int timestamp) {
  std::lock_guard<SpinMutex> lk(tracker_mu_);
  Slot* s = Find(timestamp);
  return s ? s->received : 0;
}

void Customer::CountResponse(int ts, int num) {
  bool wake = false;
  {
    std::lock_guard<SpinMutex> lk(tracker_mu_);
    Slot* s = Find(ts);
    if (s) {
      s->received += num;
      if (s->received >= s->expected) {
        completions_.fetch_add(1, std::memory_order_release);
        wake = s->waiters > 0;
      }
    }
  }
  if (wake) tracker_cv_.notify_all();
}

void Customer::AddResponse(int timestamp, int num) { CountResponse(timestamp, num); }

void Customer::Deliver(const Message& m) {
  EventTrace::Mark("handle", m.meta.timestamp, m.meta.request * 2 + m.meta.push);
  recv_handle_(m);
  if (!m.meta.request) CountResponse(m.meta.timestamp, 1);
}

void Customer::set_inline_dispatch(bool on) {
  static const bool handoff = GetEnv("PS_LOCAL_HANDOFF", 0) != 0;  // senders deliver too: not one thread
  inline_max_bytes_.store(GetEnv("PS_INLINE_MAX_BYTES", 65536), std::memory_order_relaxed);
  inline_.store(on && !handoff, std::memory_order_release);
}

bool Customer::TryInline(const Message& m) {
  if (!inline_.load(std::memory_order_acquire) || pending_.load(std::memory_order_acquire) != 0) return false;
  // bytes this thread would have to stream itself: a payload that arrived IN the frame (a one-sided
  // descriptor names bytes that are already in place, however many: Meta::mem is valid) and the
  // reply a two-sided pull asks for. A receive thread that streams megabytes into a peer's ring
  // while that peer's receive thread does the same towards us would never drain its own ring.
  const bool one_sided = m.meta.mem.valid();
  const int64_t limit = inline_max_bytes_.load(std::memory_order_relaxed);
  if (!one_sided && m.meta.data_size > limit) return false;
  if (!one_sided && m.meta.request && !m.meta.push && m.meta.val_len > limit) return false;
  if (!deliver_mu_.try_lock()) return false;  // another van's receive thread is in the handler
  if (pending_.load(std::memory_order_acquire) != 0) {  // it queued something meanwhile: keep the order
    deliver_mu_.unlock();
    return false;
  }
  Deliver(m);
  deliver_mu_.unlock();
  return true;
}

void Customer::Accept(const Message& recved) {
  if (direct_dispatch_) {
    // several vans (MultiVan rails) may deliver at once: still one handler at a time
    std::lock_guard<SpinMutex> one(deliver_mu_);
    Deliver(recved);
  } else if (TryInline(recved)) {
  } else {
    pending_.fetch_add(1, std::memory_order_acq_rel);
    inbox_.Push(recved);
  }
}

void Customer::Accept(Message&& recved) {
  if (direct_dispatch_) {
    std::lock_guard<SpinMutex> one(deliver_mu_);
    Deliver(recved);
  } else if (TryInline(recved)) {
  } else {
    pending_.fetch_add(1, std::memory_order_acq_rel);
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
      {
        std::lock_guard<SpinMutex> one(deliver_mu_);
        Deliver(m);
      }
      pending_.fetch_sub(1, std::memory_order_acq_rel);
      continue;
    }
    bool stop = false;
    {
      Van::CorkScope cork(postoffice_->van());
      std::lock_guard<SpinMutex> one(deliver_mu_);
      Deliver(m);
      int handled = 1;
      for (int n = 1; n < 32 && inbox_.TryPop(&m); ++n) {
        if (!m.meta.control.empty() && m.meta.control.cmd == Control::TERMINATE) {
          stop = true;
          break;
        }
        Deliver(m);
        ++handled;
      }
      pending_.fetch_sub(handled, std::memory_order_acq_rel);
    }
    if (stop) break;
  }
}

}  // namespace ps
