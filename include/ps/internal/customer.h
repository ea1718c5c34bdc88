/**
 * \file customer.h
 * \brief Per-(app, customer) inbox and request tracker.
 *
 * A Customer owns one inbox (filled by the van receive thread through Accept)
 * and a tracker that counts responses per request timestamp so that Wait(ts)
 * can block until every addressed server has answered.
 *
 * Re-designed relative to the reference (include/ps/internal/customer.h:28-115,
 * src/customer.cc:18-74):
 *  - the tracker is a power-of-two ring recycled by `ts & mask` (the reference's
 *    vector grows by one entry per request forever);
 *  - PS_DIRECT_DISPATCH=1 runs the receive handle inline on the van thread
 *    (one hop instead of two) for handlers that never block, e.g. the GPU server
 *    engine which only enqueues kernels.
 */
#ifndef PS_INTERNAL_CUSTOMER_H_
#define PS_INTERNAL_CUSTOMER_H_
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>
#include "ps/internal/message.h"
#include "ps/internal/spin_mutex.h"
#include "ps/internal/threadsafe_queue.h"

namespace ps {

class Postoffice;

class Customer {
 public:
  /*! \brief invoked (on the customer thread, or inline) for every message addressed here */
  typedef std::function<void(const Message& recved)> RecvHandle;

  /*!
   * \param start_now register with the postoffice and start receiving inside the constructor.
   *        Owners that store the pointer in a member their handler uses (KVWorker / KVServer /
   *        SimpleApp keep it in `obj_`) pass false and call Start() after the assignment: a
   *        message parked for this customer would otherwise reach the handler before `obj_` is set.
   */
  Customer(int app_id, int customer_id, const RecvHandle& recv_handle, Postoffice* postoffice,
           bool start_now = true);
  /*! \brief register + start the receive thread (idempotent) */
  void Start();
  /*!
   * \brief run the handler on the van's receive thread whenever this customer's queue is idle
   *        (one thread hop less per message; a hop costs microseconds when the next thread polls
   *        and 100+ us when it sleeps). Order and exclusivity are kept: a message is handled
   *        inline only while nothing is queued or being handled by the customer thread.
   *        ONLY for handlers that never wait for the network: the receive thread is the one that
   *        would deliver what they wait for, and a handler that *sends* from it must not have a
   *        peer that does the same (two receive threads blocked on each other's full ring).
   *        Messages whose payload (or requested reply) travels inside frames and exceeds
   *        PS_INLINE_MAX_BYTES (64 KB) still go through the customer thread: streaming megabytes
   *        through a ring from the receive thread would stop it from receiving meanwhile.
   *        May be switched at any time. Ignored with PS_LOCAL_HANDOFF (several delivering threads).
   */
  void set_inline_dispatch(bool on);
  bool inline_dispatch() const { return inline_.load(std::memory_order_acquire); }
  ~Customer();
  Customer(const Customer&) = delete;
  Customer& operator=(const Customer&) = delete;

  // ---- identity -------------------------------------------------------------
  int app_id() const { return app_id_; }
  int customer_id() const { return customer_id_; }

  // ---- inbox (van receive thread -> application) ------------------------------
  /*! \brief hand a received message to this customer */
  void Accept(const Message& recved);
  /*! \brief same, taking ownership: saves the deep copy of the Meta on the receive thread */
  void Accept(Message&& recved);

  // ---- request tracker -----------------------------------------------------------
  /*!
   * \brief open a request addressed to node group `recver`; returns its timestamp.
   *  The expected response count is the number of *groups* in `recver` (a worker
   *  instance talks to exactly one instance of each server group), 1 for a single
   *  node id, or `num_expected` if given.
   */
  int NewRequest(int recver, int num_expected = -1);
  /*! \brief count `num` responses that will never arrive as messages (e.g. empty slices) */
  void AddResponse(int timestamp, int num = 1);
  /*! \brief responses counted so far for `timestamp` */
  int NumResponse(int timestamp);
  /*! \brief block until `timestamp` has all its responses */
  void WaitRequest(int timestamp);

 private:
  /*! \brief one in-flight request; lives in ring_[ts & mask] until a later ts recycles it */
  struct Slot {
    int ts = -1;
    int expected = 0;
    int received = 0;
    int waiters = 0;  // threads blocked in WaitRequest(ts): only they are woken, and only on completion
  };
  Slot* Find(int ts);  // caller holds tracker_mu_
  /*! \brief count `num` responses for `ts`; wakes the request's waiters when it completes */
  void CountResponse(int ts, int num);
  void Deliver(const Message& m);
  void Receiving();

  const int app_id_;
  const int customer_id_;
  RecvHandle recv_handle_;
  Postoffice* postoffice_;
  std::atomic<uint64_t> wait_sleeps_{0};  // WaitRequest calls that outlasted the spin window
  bool direct_dispatch_ = false;       // PS_DIRECT_DISPATCH=1: no customer thread at all
  std::atomic<bool> inline_{false};    // set_inline_dispatch
  std::atomic<int> pending_{0};        // queued or being handled by the customer thread
  SpinMutex deliver_mu_;               // one handler at a time (several vans may deliver: MultiVan)
  std::atomic<int64_t> inline_max_bytes_{65536};  // in-frame payloads above this go to the customer thread
  bool TryInline(const Message& m);
  bool started_ = false;

  ThreadsafeQueue<Message> inbox_;
  std::unique_ptr<std::thread> recv_thread_;

  SpinMutex tracker_mu_;
  std::condition_variable_any tracker_cv_;
  std::vector<Slot> ring_;
  int next_ts_ = 0;
  /*! \brief requests completed so far: lets a waiter poll without taking tracker_mu_ */
  std::atomic<uint64_t> completions_{0};
};

}  // namespace ps
#endif  // PS_INTERNAL_CUSTOMER_H_
