/**
 * \file van.h
 * \brief Van: the transport-independent half of the communication stack.
 *
 * The base class owns the control protocol (node registration and id
 * assignment at the scheduler, group / instance barriers, heartbeats,
 * terminate, optional at-least-once resending, receive-side fault injection,
 * van-level profiling) and the receive loop. A transport subclass supplies
 * Connect / Bind / SendMsg / RecvMsg. Available transports (Van::Create):
 *   "zmq" | "0" | "tcp"   native epoll TCP / unix-socket van (control + CPU data)
 *   "multivan"            N inner tcp vans, one per port / device
 *   "shm"                 one-sided protocol over POSIX shared memory (CPU)
 *   "nvl" | "1" | "ibverbs" | "ucx"
 *                         one-sided protocol over NVLink peer memory (B200)
 * API parity: reference include/ps/internal/van.h:29-254, src/van.cc.
 */
#ifndef PS_INTERNAL_VAN_H_
#define PS_INTERNAL_VAN_H_
#include <atomic>
#include "ps/internal/symmetric.h"
#include <ctime>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include "ps/base.h"
#include "ps/internal/message.h"

namespace ps {

class Resender;
class Postoffice;

class Van {
 public:
  /*! \brief factory; `type` is the value of DMLC_ENABLE_RDMA (see file comment) */
  static Van* Create(const std::string& type, Postoffice* postoffice);

  explicit Van(Postoffice* postoffice);
  virtual ~Van();

  /*!
   * \brief bind, connect to the scheduler, register, wait until the cluster is
   *        assembled. With `standalone` the van only binds and starts receiving
   *        (used when it is the inner channel of a composite van).
   */
  virtual void Start(int customer_id, bool standalone);
  /*! \brief send; thread-safe; returns bytes sent (>0) */
  int Send(Message& msg);
  /*! \brief like Send but a transport failure is reported (-1), not fatal */
  int SendBestEffort(Message& msg);
  const Node& my_node() const {
    CHECK(ready_.load() || my_node_set_) << "call Start() first";
    return my_node_;
  }
  /*! \brief stop the receive thread and the transport */
  virtual void Stop();
  /*! \brief next control-plane timestamp */
  int GetTimestamp() {
    int ts = timestamp_++;
    if (ts == Meta::kEmpty) ts = timestamp_++;  // the sentinel is never a real timestamp
    return ts;
  }
  bool IsReady() { return ready_.load(); }
  virtual std::string GetType() const = 0;

  /*!
   * \brief register a landing buffer for pushes of (msg.meta.sender, msg.meta.key):
   *        data[1] of later matching pushes is delivered *in that buffer*.
   */
  virtual void RegisterRecvBuffer(Message& /*msg*/) {}
  /*! \brief make [addr, addr+length) reachable by peers (export / map) ahead of use */
  virtual void PinMemory(void* /*addr*/, size_t /*length*/, bool /*gpu*/, int /*dev_index*/ = 0) {}
  /*!
   * \brief forget everything cached about the exported allocation that contains `addr` — call it
   *        BEFORE freeing memory that was pinned or used as a one-sided destination, so that a later
   *        allocation at the same address is exported afresh (new handle, new region id) instead of
   *        being mistaken for the old one.
   */
  virtual void UnpinMemory(void* /*addr*/) {}
  /*! \brief memory peers can map (HBM on the NVLink van, shm on the shm van, else heap) */
  virtual void* AllocExportable(size_t bytes) { return malloc(bytes); }
  virtual void FreeExportable(void* p) { free(p); }
  /*! \brief ... on a given device of a van whose process drives several GPUs (DMLC_NUM_GPU_DEV) */
  virtual void* AllocExportableOn(size_t bytes, int /*device*/) { return AllocExportable(bytes); }
  /*! \brief devices this van's process drives (1 unless DMLC_NUM_GPU_DEV / PS_NUM_GPU_DEV says more) */
  virtual int NumDevices() { return 1; }
  /*! \brief this process's mapping of a span a peer announced (one-sided vans), else null */
  virtual void* ResolvePeerMem(int /*node_id*/, const MemRef& /*mem*/) { return nullptr; }
  /*!
   * \brief collective over all worker and server PROCESSES of the job (one call per process and
   *        `tag`; a joint process's second van gets the first one's result): allocate `bytes` of
   *        zero-filled symmetric memory — a block in every process, each mapped by all the
   *        others, plus an NVSwitch multicast address over all of them where the hardware has
   *        one (out->mc). HBM on the nvl van, shared memory on the shm van; false elsewhere.
   */
  virtual bool AllocSymmetric(const std::string& /*tag*/, size_t /*bytes*/, SymmetricBuffer* /*out*/) { return false; }
  /*! \brief every node of the job as the scheduler announced it (empty before ADD_NODE completes) */
  std::vector<Node> ClusterNodes() {
    std::lock_guard<std::mutex> lk(cluster_mu_);
    return cluster_;
  }
  /*! \brief transport-specific counters by name (one-sided copies, gated descriptors, ...) */
  virtual void TransportStats(std::vector<std::pair<std::string, uint64_t>>* /*out*/) {}
  /*! \brief stream the van's copy kernels run on (cudaStream_t), null for CPU vans */
  virtual void* DataStream() { return nullptr; }
  /*!
   * \brief launch coalescing. Between Cork() and the matching Uncork() the calling thread's
   *        data messages are held back; Uncork() issues all their one-sided copies as one batch
   *        (few kernel launches, one completion event) and then sends the messages in the
   *        order they were submitted. Nests; other threads are unaffected. No-ops for vans
   *        that have nothing to merge.
   */
  /*!
   * \brief in-process hand-off: take a data message another van of THIS process built for us,
   *        exactly as if the receive loop had read it (byte count, verbose log, customer
   *        lookup). False — and `*msg` untouched — when per-message receive processing is
   *        active (resend bookkeeping, drop injection, profiling log), the van is not running,
   *        or the customer does not exist yet: the caller then uses the normal transport.
   */
  bool AcceptHandoff(Message* msg);
  virtual void Cork() {}
  virtual void Uncork() {}
  /*! \brief RAII helper for Cork / Uncork */
  class CorkScope {
   public:
    explicit CorkScope(Van* van) : van_(van) {
      if (van_) van_->Cork();
    }
    ~CorkScope() {
      if (van_) van_->Uncork();
    }
    CorkScope(const CorkScope&) = delete;
    CorkScope& operator=(const CorkScope&) = delete;

   private:
    Van* van_;
  };
  /*! \brief install the node identity (after the scheduler assigned the id) */
  virtual void SetNode(const Node& node) {
    my_node_ = node;
    my_node_set_ = true;
  }

  /*! \brief hand messages that arrived before their customer existed to it (called by Postoffice) */
  void DeliverParked();

  /*! \brief cumulative payload+meta byte counters */
  size_t send_bytes() const { return send_bytes_.load(); }
  size_t recv_bytes() const { return recv_bytes_.load(); }

  /*! \brief meta codec exposed for transports and tests */
  static void PackMeta(const Meta& meta, std::vector<char>* buf);
  static bool UnpackMeta(const char* buf, size_t size, Meta* meta);

 protected:
  /*! \brief open the send path to `node` */
  virtual void Connect(const Node& node) = 0;
  /*! \brief bind the receive endpoint; returns the bound port or -1 */
  virtual int Bind(Node& node, int max_retry) = 0;
  /*! \brief block for one message; returns bytes received or -1 */
  virtual int RecvMsg(Message* msg) = 0;
  /*! \brief transmit; returns bytes sent or -1 */
  virtual int SendMsg(Message& msg) = 0;

  /*! \brief hook: a control message the base class does not know (e.g. ADDR_REQUEST) */
  virtual bool HandleTransportControl(Message* /*msg*/) { return false; }

  Node scheduler_;
  Node my_node_;
  bool my_node_set_ = false;
  bool is_scheduler_ = false;
  std::mutex start_mu_;
  Postoffice* postoffice_;

  /*! \brief deliver to the application layer (customers) */
  void ProcessDataMsg(Message* msg);

 private:
  friend class Resender;
  void Receiving();
  void HeartbeatLoop();
  void OnTerminate();
  void OnAddNode(Message* msg);
  void OnBarrier(Message* msg);
  void OnInstanceBarrier(Message* msg);
  void OnHeartbeat(Message* msg);

  // -- scheduler-side registration book ------------------------------------
  void SchedulerCollect(Message* msg);
  void SchedulerAssignAndBroadcast();
  void SchedulerHandleRecovery(const Node& reborn);
  void OrderRegistrants(std::vector<Node>* nodes);
  void AdoptIdentity(const std::vector<Node>& nodes);

  std::vector<Node> registrants_;   // scheduler: nodes seen so far / final table
  bool table_final_ = false;        // scheduler: ids were assigned
  int num_servers_ = 0;
  int num_workers_ = 0;
  /*! \brief "host:port" -> node id, for every peer Connect() was issued to */
  std::unordered_map<std::string, int> connected_;
  /*! \brief ids aliased onto another id because they share a host:port */
  std::unordered_map<int, int> alias_of_;

  std::unique_ptr<std::thread> receiver_thread_;
  std::unique_ptr<std::thread> heartbeat_thread_;
  std::atomic<bool> ready_{false};
  std::atomic<bool> stopping_{false};
  std::atomic<size_t> send_bytes_{0};
  std::atomic<size_t> recv_bytes_{0};
  std::atomic<int> timestamp_{0};
  int init_stage_ = 0;
  int heartbeat_timeout_ = 0;
  Resender* resender_ = nullptr;
  int drop_rate_ = 0;
  unsigned drop_seed_ = 0;

  std::mutex parked_mu_;
  std::vector<Message> parked_;  // data messages waiting for their customer to be created
  std::atomic<int> parked_n_{0};  // parked_.size(), readable without the lock
  std::mutex cluster_mu_;
  std::vector<Node> cluster_;     // node table of the job (workers and servers), by arrival
  void DrainParkedLocked();

  std::vector<int> instance_barrier_count_;
  std::unordered_map<int, std::vector<int>> group_barrier_requests_;

  // van-level profiling (ENABLE_PROFILING=1); one stream per van instance
  bool profiling_ = false;
  std::ofstream profile_out_;
  std::mutex profile_mu_;

  Van(const Van&) = delete;
  Van& operator=(const Van&) = delete;
};

}  // namespace ps
#endif  // PS_INTERNAL_VAN_H_
