/**
 * \file tcp_van.h
 * \brief TcpVan: the CPU / control-plane transport (answers to "zmq", "0", "tcp").
 *
 * Fills the role of the reference's ZMQVan (src/zmq_van.h:43-480) without
 * libzmq: plain stream sockets, one outbound connection per peer, one epoll
 * loop for everything inbound, and a loopback queue for self-addressed
 * messages. Design points:
 *   - frame = [FrameHeader][u64 seg_len x n][meta][seg0][seg1]... sent with a
 *     single gathered sendmsg(); payload SArrays are never copied on send.
 *   - RecvMsg runs the epoll loop *in the van receive thread itself* — there is
 *     no transport-side thread or intermediate queue (the reference has one
 *     recv thread per socket pushing into a queue while holding the send mutex).
 *   - payload bytes are read straight into their final buffer: a registered
 *     receive buffer if one matches (sender, key), else a fresh page-aligned one.
 *   - DMLC_LOCAL=1 switches to abstract-namespace unix sockets ("ipc" mode).
 */
#ifndef PS_VAN_TCP_VAN_H_
#define PS_VAN_TCP_VAN_H_
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <set>
#include <shared_mutex>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include "ps/internal/postoffice.h"
#include "ps/internal/van.h"
#include "van/mem_domain.h"
#include "core/event_trace.h"
#include "ps/internal/spin_mutex.h"
#include "van/shm_pipe.h"
#include "van/shm_util.h"

namespace ps {

/*!
 * \brief recycles page-aligned receive buffers by size so steady-state traffic never
 *        page-faults fresh memory (the dominant cost of a 1 MB receive on loopback)
 */
class RecvBufferPool : public std::enable_shared_from_this<RecvBufferPool> {
 public:
  explicit RecvBufferPool(size_t cap_bytes) : cap_(cap_bytes) {}
  ~RecvBufferPool() {
    for (auto& kv : free_)
      for (char* p : kv.second) free(p);
  }
  SArray<char> Get(size_t len) {
    const size_t rounded = (len + 4095) & ~size_t(4095);
    char* p = nullptr;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = free_.find(rounded);
      if (it != free_.end() && !it->second.empty()) {
        p = it->second.back();
        it->second.pop_back();
        cached_ -= rounded;
      }
    }
    if (!p) {
      void* vp = nullptr;
      CHECK_EQ(posix_memalign(&vp, 4096, rounded), 0);
      p = static_cast<char*>(vp);
    }
    std::shared_ptr<RecvBufferPool> self = shared_from_this();
    SArray<char> seg;
    seg.reset(p, len, [self, rounded](char* q) { self->Put(q, rounded); });
    return seg;
  }

 private:
  void Put(char* p, size_t rounded) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (cached_ + rounded <= cap_) {
        free_[rounded].push_back(p);
        cached_ += rounded;
        return;
      }
    }
    free(p);
  }
  std::mutex mu_;
  size_t cap_;
  size_t cached_ = 0;
  std::map<size_t, std::vector<char*>> free_;
};

class TcpVan : public Van {
 public:
  explicit TcpVan(Postoffice* postoffice) : Van(postoffice) {}
  ~TcpVan() override {
    StopOutbox();
    CloseAll();
  }

  std::string GetType() const override { return "zmq"; }

  void Start(int customer_id, bool standalone) override {
    InitTransport();
    Van::Start(customer_id, standalone);
    if (handoff_ && AllowLocalHandoff() && !standalone) {
      std::unique_lock<std::shared_mutex> lk(LocalVans().mu);
      LocalVans().by_port[my_node_.port] = this;
    }
  }

  void Stop() override {
    {
      std::unique_lock<std::shared_mutex> lk(LocalVans().mu);  // waits for hand-offs in flight
      auto& m = LocalVans().by_port;
      for (auto it = m.begin(); it != m.end();) it = it->second == this ? m.erase(it) : std::next(it);
    }
    Van::Stop();
    StopOutbox();
    CloseAll();
  }

  /*! \brief later pushes of (msg.meta.sender, msg.meta.key) land in msg.data[1] */
  void RegisterRecvBuffer(Message& msg) override {
    CHECK_GE(msg.data.size(), (size_t)2);
    std::lock_guard<std::mutex> lk(reg_mu_);
    registered_[std::make_pair(msg.meta.sender, msg.meta.key)] = msg.data[1];
  }

 protected:
  /*! \brief create the epoll set / wake fd / buffer pool (idempotent) */
  void InitTransport() {
    std::lock_guard<std::mutex> lk(init_mu_);
    if (epfd_ >= 0) return;
    local_ipc_ = GetEnv("DMLC_LOCAL", 0) != 0;
    connect_timeout_s_ = GetEnv("PS_CONNECT_TIMEOUT", 120);
    // landing a pull reply straight in the caller's buffer is only safe if every reply is
    // wanted: with PS_RESEND a retransmitted duplicate can arrive after the caller has
    // already consumed the first copy and released the buffer
    direct_pull_ = GetEnv("PS_TCP_DIRECT_PULL", 1) != 0 && GetEnv("PS_RESEND", 0) == 0;
    // BYTEPS_ENABLE_IPC=0 (the reference's switch for its shared-memory side transport,
    // src/rdma_van.h:44-46) turns the same-host rings off as well
    use_pipes_ = GetEnv("PS_SHM_PIPE", 1) != 0 && GetEnv("BYTEPS_ENABLE_IPC", 1) != 0;
    pipe_bytes_ = static_cast<size_t>(GetEnv("PS_SHM_PIPE_KB", 256)) << 10;
    if (use_pipes_) {
      static const int swept = SweepStaleShm("pslb200_");  // rings of processes that were killed
      (void)swept;
    }
    if (!pool_) {
      pool_ = std::make_shared<RecvBufferPool>(static_cast<size_t>(GetEnv("PS_TCP_POOL_MB", 1024))
                                               << 20);
    }
    epfd_ = epoll_create1(EPOLL_CLOEXEC);
    CHECK_GE(epfd_, 0) << strerror(errno);
    wake_fd_ = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
    CHECK_GE(wake_fd_, 0) << strerror(errno);
    AddToEpoll(wake_fd_);
  }

  // the wire prefix of every frame
  struct FrameHeader {
    uint32_t magic;
    int32_t sender;
    int32_t recver;
    uint32_t meta_len;
    uint32_t num_segments;
    uint32_t reserved;
    /*! \brief 0, or "deliver only once the ring's completion word has reached this value" */
    uint64_t gate;
  };
  static constexpr uint32_t kFrameMagic = 0x4d524650u;  // "PFRM"
  static constexpr uint32_t kMaxSegments = 16;

  struct Peer {
    int fd = -1;
    std::mutex mu;  // serialises whole frames on this socket
    /*! \brief same-host fast path: frames go through this ring, the socket carries doorbells */
    std::unique_ptr<ShmPipe> pipe;
    /*! \brief a ring the peer has not accepted yet (and the gate mapping that goes with it) */
    std::unique_ptr<ShmPipe> offered;
    void* offered_gate_word = nullptr;
    /*! \brief the peer lives in this very process (joint roles, in-process clusters) */
    bool same_process = false;
    int port = 0;
    /*! \brief hand-off decision for data messages: 0 untried, 1 direct, 2 wire only. A peer that
     *  ever needed the wire keeps using it, so a handed-off message can never overtake one
     *  that is still in the ring (or parked at the receiver). */
    std::atomic<int> handoff_state{0};
    /*! \brief gated frames (see ShmPipe): sequence of the last one, the address the copy engine
     *  stores completions to (null: no gating on this connection), and what must stay alive
     *  until a given completion. All under `mu`. */
    /*! \brief messages of receive threads parked in the outbox for this peer (see Defer) */
    std::atomic<int> deferred{0};
    /*! \brief we offered this peer a ring (so it will offer us one too) */
    bool made_offer = false;
    uint64_t gate_seq = 0;
    void* gate_word = nullptr;
    std::deque<std::pair<uint64_t, SArray<char>>> gate_keep;
  };
  /*!
   * \brief issues the work a gated frame announces. Called under the peer's lock, after the
   *        frame's sequence number is fixed and BEFORE the frame is published, so the order of
   *        completions on `word` is the order of the frames in the ring. Must arrange for `seq`
   *        to be stored to `word` when the work is globally visible; false = nothing was issued.
   */
  using GateIssue = std::function<bool(void* word, uint64_t seq)>;
  /*! \brief transient MemRef::region marker, never on the wire (see SendMsg) */
  static constexpr int32_t kEncodedOnHost = 0x4000007f;
  // Same-host ring negotiation (all on the connection's socket):
  //   sender   -> receiver : kPipeMagic + name      "I created shm ring <name>, can you map it?"
  //   receiver -> sender   : one byte 'A' / 'D'     accept (mapped) or decline (e.g. a container
  //                                                 with the same IP but a private /dev/shm)
  //   sender   -> receiver : kPipeSwitchMagic       "every later frame is in the ring"
  // Until the switch the socket carries the frames, so a declined or unanswered offer costs nothing.
  static constexpr uint32_t kPipeMagic = 0x45504950u;        // "PIPE"
  static constexpr uint32_t kPipeSwitchMagic = 0x57535050u;  // "PPSW"

  int Bind(Node& node, int max_retry) override {
    int port = node.port;
    unsigned seed = static_cast<unsigned>(time(nullptr)) + static_cast<unsigned>(getpid()) +
                    static_cast<unsigned>(port);
    int fixed_port_waits = 0;
    for (int attempt = 0; attempt <= max_retry; ++attempt) {
      int fd = -1;
      bool ok = false;
      if (local_ipc_) {
        fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
        CHECK_GE(fd, 0) << strerror(errno);
        struct sockaddr_un sa;
        socklen_t len = UnixAddr(port, &sa);
        ok = bind(fd, reinterpret_cast<struct sockaddr*>(&sa), len) == 0;
      } else {
        fd = socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
        CHECK_GE(fd, 0) << strerror(errno);
        int one = 1;
        setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
        struct sockaddr_in sa;
        memset(&sa, 0, sizeof(sa));
        sa.sin_family = AF_INET;
        sa.sin_addr.s_addr = htonl(INADDR_ANY);
        sa.sin_port = htons(static_cast<uint16_t>(port));
        ok = bind(fd, reinterpret_cast<struct sockaddr*>(&sa), sizeof(sa)) == 0;
      }
      if (ok && listen(fd, 1024) == 0) {
        listen_fd_ = fd;
        AddToEpoll(fd);
        return port;
      }
      const int why = errno;
      close(fd);
      if (max_retry == 0 && why == EADDRINUSE && fixed_port_waits < 30) {
        // a fixed port (the scheduler's) that something — typically a short-lived outgoing
        // connection that was handed this number — still occupies: wait for it, do not give up
        ++fixed_port_waits;
        --attempt;
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
        continue;
      }
      if (attempt == max_retry) break;
      port = 10000 + static_cast<int>(rand_r(&seed) % 40000);
    }
    return -1;
  }

  void Connect(const Node& node) override {
    CHECK_NE(node.id, Node::kEmpty);
    CHECK_NE(node.port, Node::kEmpty);
    CHECK(!node.hostname.empty());
    // same-role peers never exchange messages (except a node with itself,
    // which is served by the loopback queue and needs no socket)
    if (node.id == my_node_.id) return;
    if (node.role == my_node_.role && node.role != Node::SCHEDULER) return;

    int fd = -1;
    const auto deadline =
        std::chrono::steady_clock::now() + std::chrono::seconds(connect_timeout_s_);
    for (;;) {
      fd = local_ipc_ ? ConnectUnix(node.port) : ConnectTcp(node.hostname, node.port);
      if (fd >= 0) break;
      CHECK(std::chrono::steady_clock::now() < deadline)
          << "cannot connect to " << node.DebugString() << ": " << strerror(errno);
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
    std::shared_ptr<Peer> peer(new Peer());
    peer->fd = fd;
    peer->port = node.port;
    peer->same_process = node.pid != 0 && node.pid == static_cast<int>(getpid()) &&
                         node.hostname == my_node_.hostname;
    if (use_pipes_ && !my_node_.hostname.empty() && node.hostname == my_node_.hostname) {
      OfferPipe(peer.get(), node.id);
    }
    std::shared_ptr<Peer> old;
    {
      std::lock_guard<SpinMutex> lk(peers_mu_);
      auto it = peers_.find(node.id);
      if (it != peers_.end()) old = it->second;
      peers_[node.id] = peer;
    }
    if (old) {  // reconnect after recovery: retire the stale socket
      std::lock_guard<std::mutex> lk(old->mu);
      if (old->fd >= 0) close(old->fd);
      old->fd = -1;
      if (old->pipe && old->gate_word) {
        ReleaseGateWord(old->pipe.get());
        old->gate_word = nullptr;
        old->gate_keep.clear();
      }
      if (old->offered && old->offered_gate_word) {
        ReleaseGateWord(old->offered.get());
        old->offered_gate_word = nullptr;
      }
    }
  }

  int SendMsg(Message& msg) override {
    CHECK_NE(msg.meta.recver, Meta::kEmpty);
    // a push that asks for a wire codec but travels two-sided (host values a one-sided van
    // could not export, or the plain TCP van): encode here, the receiver sees wire bytes
    if (msg.meta.request && msg.meta.push && msg.meta.codec > kCodecRaw && msg.meta.codec < kCodecPlaced &&
        msg.data.size() >= 2 && msg.data[1].size() > 0 && !msg.data[1].on_gpu() && !msg.meta.mem.valid()) {
      const size_t raw = msg.data[1].size();
      SArray<char> encoded(WireBytes(msg.meta.codec, raw));
      CHECK_EQ(ps_host_copy(encoded.data(), msg.data[1].data(), raw, msg.meta.codec, msg.meta.scale), 0);
      Message wire_msg = msg;
      wire_msg.data[1] = encoded;
      wire_msg.meta.data_size += static_cast<int64_t>(encoded.size()) - static_cast<int64_t>(raw);
      wire_msg.meta.mem.region = kEncodedOnHost;  // "do not encode again", cleared by SendFrame
      return SendFrame(wire_msg);
    }
    return SendFrame(msg);
  }

  /*! \brief how often the receive thread gave up polling and slept (each wake-up costs 50-300 us on a VM) */
  uint64_t num_blocking_waits() const { return blocking_waits_.load(); }
  /*! \brief sends of receive threads that went through the outbox because they would have had to wait */
  uint64_t num_deferred_sends() const { return deferred_sends_.load(); }

  /*! \brief can frames to `recver` be gated on completions the copy engine signals itself? */
  bool PeerGated(int recver) {
    std::lock_guard<SpinMutex> lk(peers_mu_);
    auto it = peers_.find(recver);
    return it != peers_.end() && it->second->pipe && it->second->gate_word != nullptr;
  }

  /*!
   * \brief ring negotiation with a same-host peer reached a verdict: it answered OUR offer, or we
   *        answered ITS offer (`accepted`: the ring could be mapped). A declined ring means the two
   *        processes share an address but not /dev/shm (containers): one-sided vans then treat
   *        the peer like one on another host. Called on the receive thread, no locks held.
   */
  virtual void OnPipeVerdict(int /*peer_id*/, bool /*accepted*/) {}

  /*!
   * \brief is the ring negotiation with same-host peer `id` still open? True while our offer is
   *        unanswered, or — if we offered one — while the peer's own offer has not arrived yet (both
   *        sides offer when both have rings enabled).
   */
  bool NegotiationOpen(int id) {
    std::shared_ptr<Peer> peer;
    {
      std::lock_guard<SpinMutex> lk(peers_mu_);
      auto it = peers_.find(id);
      if (it == peers_.end()) return false;
      peer = it->second;
    }
    bool offered_one = false;
    {
      std::lock_guard<std::mutex> lk(peer->mu);
      if (peer->offered) return true;
      offered_one = peer->made_offer;
    }
    if (!offered_one) return false;
    std::lock_guard<std::mutex> lk(offer_mu_);
    return offers_seen_.count(id) == 0;
  }

  /*! \brief is the caller this van's receive thread? (it must not wait for events only it can process) */
  bool OnReceiveThread() const { return tls_receiving_ == this; }

  /*! \brief is there a connection to node `id` already (then Connect(id) is a reconnect after recovery) */
  bool HasPeer(int id) {
    std::lock_guard<SpinMutex> lk(peers_mu_);
    return peers_.count(id) > 0;
  }

  /*! \brief -2 from SendFrame: the gate could not be issued, nothing was sent */
  static constexpr int kNotGated = -2;

  /*!
   * \brief serialise one message to its peer (socket, ring or own loopback queue). With `issue`
   *        the frame is gated: it enters the ring now but the receiver delivers it only after the
   *        work `issue` started has signalled completion; `keep` stays referenced until then.
   */
  int SendFrame(Message& msg, const GateIssue* issue = nullptr, const SArray<char>* keep = nullptr) {
    EventTrace::Mark("send_frame", msg.meta.timestamp, msg.meta.request * 2 + msg.meta.push);
    if (msg.meta.mem.region == kEncodedOnHost) msg.meta.mem = MemRef();
    const int recver = msg.meta.recver;
    if (direct_pull_ && msg.meta.request && msg.meta.control.empty() && msg.meta.src_dev_type != GPU) {
      // remember where the reply of this pull may land: the address echoed back on the wire is
      // only ever compared with this record, never trusted (see SegmentDestination). A request
      // that DESCRIBES its destination (one-sided vans) is answered by a write into it and a
      // descriptor: no segment will ever arrive to consume a record, so none is made.
      const bool fused = msg.meta.push && msg.meta.pull;
      const bool described = fused ? msg.meta.pull_mem.valid() : msg.meta.mem.valid();
      const uint64_t addr = described ? 0 : (fused ? msg.meta.pull_addr : (msg.meta.push ? 0 : msg.meta.addr));
      if (addr != 0) {
        const size_t esz = msg.meta.data_type.size() > 1 ? ElemSize(msg.meta.data_type[1]) : 1;
        const uint64_t bytes = static_cast<uint64_t>(fused ? msg.meta.pull_len : msg.meta.val_len) * esz;
        std::lock_guard<SpinMutex> lk(pull_mu_);
        // (replies that never come — a dead server, an empty value — must not pile up for ever)
        if (pull_dests_.size() >= kMaxPullRecords) pull_dests_.erase(pull_dests_.begin());
        pull_dests_[PullKey(recver, msg.meta.app_id, msg.meta.customer_id, msg.meta.timestamp)] = {addr, bytes};
      }
    }
    if (recver == my_node_.id) return Loopback(msg);

    std::shared_ptr<Peer> peer;
    {
      std::lock_guard<SpinMutex> lk(peers_mu_);
      auto it = peers_.find(recver);
      if (it == peers_.end()) {
        LOG(WARNING) << "there is no socket to node " << recver;
        return -1;
      }
      peer = it->second;
    }
    const bool transport_ctrl = msg.meta.control.cmd == Control::ADDR_REQUEST ||
                                msg.meta.control.cmd == Control::ADDR_RESOLVED;
    // a gated frame needs the ring (its completion word lives there); from then on everything
    // for this peer does, so nothing handed off later can overtake it
    if (issue) peer->handoff_state.store(2, std::memory_order_release);
    if (peer->same_process && handoff_ && (msg.meta.control.empty() || transport_ctrl) &&
        peer->handoff_state.load(std::memory_order_acquire) != 2) {
      if (HandOff(peer.get(), msg)) {
        peer->handoff_state.store(1, std::memory_order_release);
        return 1 + static_cast<int>(msg.meta.data_size & 0x3fffffff);
      }
      peer->handoff_state.store(2, std::memory_order_release);
    }
    thread_local std::vector<char> meta_buf;  // reused: no allocation per message
    meta_buf.clear();
    PackMeta(msg.meta, &meta_buf);
    const uint32_t nseg = static_cast<uint32_t>(msg.data.size());
    CHECK_LE(nseg, kMaxSegments);
    FrameHeader hdr;
    hdr.magic = kFrameMagic;
    hdr.sender = my_node_.id;
    hdr.recver = recver;
    hdr.meta_len = static_cast<uint32_t>(meta_buf.size());
    hdr.num_segments = nseg;
    hdr.reserved = 0;
    hdr.gate = 0;
    uint64_t seg_len[kMaxSegments];
    struct iovec iov[3 + kMaxSegments];
    int niov = 0;
    size_t total = 0;
    iov[niov++] = {&hdr, sizeof(hdr)};
    if (nseg) iov[niov++] = {seg_len, sizeof(uint64_t) * nseg};
    iov[niov++] = {meta_buf.data(), meta_buf.size()};
    for (uint32_t i = 0; i < nseg; ++i) {
      seg_len[i] = msg.data[i].size();
      CHECK(seg_len[i] == 0 || !msg.data[i].on_gpu())
          << "TcpVan cannot send device memory: " << msg.DebugString();
      if (seg_len[i]) iov[niov++] = {msg.data[i].data(), msg.data[i].size()};
    }
    for (int i = 0; i < niov; ++i) total += iov[i].iov_len;

    std::unique_lock<std::mutex> lk(peer->mu, std::defer_lock);
    if (tls_receiving_ != nullptr && !tls_outbox_) {
      // A receive thread must never wait for a connection: another thread of this process may be
      // streaming megabytes into it (holding the lock, or having filled the ring), which only ends
      // when the PEER's receive thread drains them — and that one may be waiting, symmetrically, for
      // us. Replies, ACKs and rendezvous messages sent from a receive thread therefore go out
      // through the outbox thread whenever they would have to wait; per peer they keep their order.
      if (issue) {
        while (peer->deferred.load(std::memory_order_acquire) != 0) std::this_thread::yield();
        lk.lock();  // gated frames are descriptors: nobody holds this lock for long on a one-sided van
      } else {
        bool wait_needed = peer->deferred.load(std::memory_order_acquire) != 0 || !lk.try_lock();
        if (!wait_needed && peer->pipe && total <= peer->pipe->capacity() / 2 && peer->pipe->FreeSpace() < total) {
          lk.unlock();
          wait_needed = true;
        }
        if (wait_needed) {
          Defer(peer, msg);
          return static_cast<int>(std::min<size_t>(total, 0x7fffffff));
        }
      }
    } else {
      lk.lock();
    }
    if (peer->fd < 0) return -1;
    if (issue) {
      if (!peer->pipe || !peer->gate_word) return kNotGated;
      // completions seen so far release what they kept alive
      const uint64_t done = peer->pipe->gate_done();
      while (!peer->gate_keep.empty() && peer->gate_keep.front().first <= done) peer->gate_keep.pop_front();
      const uint64_t seq = peer->gate_seq + 1;
      if (!(*issue)(peer->gate_word, seq)) return kNotGated;
      peer->gate_seq = seq;
      hdr.gate = seq;
      if (keep && keep->size()) peer->gate_keep.emplace_back(seq, *keep);
    }
    if (!(peer->pipe ? SendThroughPipe(peer.get(), iov, niov) : SendAll(peer->fd, iov, niov))) {
      LOG(WARNING) << "failed to send to node " << recver << ": " << strerror(errno);
      return -1;
    }
    return static_cast<int>(std::min<size_t>(total, 0x7fffffff));
  }

  /*!
   * \brief several messages for ONE peer in a single gathered sendmsg (frames back to back).
   *        Used by the one-sided van to flush a backlog of descriptors with one system call.
   */
  int SendMsgBatch(int recver, const std::vector<Message*>& msgs) {
    if (msgs.empty()) return 0;
    if (msgs.size() == 1 || recver == my_node_.id) {
      int total = 0;
      for (Message* m : msgs) {
        // qualified on purpose: a derived van's SendMsg would treat the descriptor as a
        // fresh message (and strip the MemRef it carries)
        int n = TcpVan::SendMsg(*m);
        if (n < 0) return -1;
        total += n;
      }
      return total;
    }
    std::shared_ptr<Peer> peer;
    {
      std::lock_guard<SpinMutex> lk(peers_mu_);
      auto it = peers_.find(recver);
      if (it == peers_.end()) return -1;
      peer = it->second;
    }
    const size_t n = msgs.size();
    std::vector<FrameHeader> hdrs(n);
    std::vector<std::array<uint64_t, kMaxSegments>> lens(n);
    std::vector<std::vector<char>> metas(n);
    std::vector<struct iovec> iov;
    iov.reserve(n * 6);
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) {
      Message& m = *msgs[i];
      PackMeta(m.meta, &metas[i]);
      const uint32_t nseg = static_cast<uint32_t>(m.data.size());
      CHECK_LE(nseg, kMaxSegments);
      hdrs[i] = {kFrameMagic, my_node_.id, recver, static_cast<uint32_t>(metas[i].size()), nseg, 0, 0};
      iov.push_back({&hdrs[i], sizeof(FrameHeader)});
      if (nseg) iov.push_back({lens[i].data(), sizeof(uint64_t) * nseg});
      iov.push_back({metas[i].data(), metas[i].size()});
      for (uint32_t sgi = 0; sgi < nseg; ++sgi) {
        lens[i][sgi] = m.data[sgi].size();
        CHECK(lens[i][sgi] == 0 || !m.data[sgi].on_gpu()) << "TcpVan cannot send device memory";
        if (lens[i][sgi]) iov.push_back({m.data[sgi].data(), m.data[sgi].size()});
      }
    }
    for (auto& v : iov) total += v.iov_len;
    std::lock_guard<std::mutex> lk(peer->mu);
    if (peer->fd < 0) return -1;
    if (peer->pipe) {
      if (!SendThroughPipe(peer.get(), iov.data(), static_cast<int>(iov.size()))) return -1;
      return static_cast<int>(std::min<size_t>(total, 0x7fffffff));
    }
    // IOV_MAX is 1024 on Linux: flush in slices
    size_t at = 0;
    while (at < iov.size()) {
      const int cnt = static_cast<int>(std::min<size_t>(512, iov.size() - at));
      if (!SendAll(peer->fd, iov.data() + at, cnt)) return -1;
      at += static_cast<size_t>(cnt);
    }
    return static_cast<int>(std::min<size_t>(total, 0x7fffffff));
  }

  int RecvMsg(Message* msg) override {
    tls_receiving_ = this;  // from here on this thread's sends must not wait for a connection (see SendFrame)
    // busy-poll window of this call: as long as the recent gaps between bursts suggest
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = RecvOne(msg, pipe_budget_.window_us());
    pipe_budget_.Observe(
        std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count());
    return rc;
  }

  int RecvOne(Message* msg, int spin_window_us) {
    msg->data.clear();
    auto last_activity = std::chrono::steady_clock::now();
    for (;;) {
      gates_pending_ = false;
      if (PollDeferred(msg)) return 1 + static_cast<int>(msg->meta.data_size & 0x3fffffff);
      if (PopLoopback(msg)) return 1 + static_cast<int>(msg->meta.data_size & 0x3fffffff);
      // same-host peers: frames arrive in shared-memory rings (round-robin for fairness)
      if (int bytes = PollPipes(msg)) {
        EventTrace::Mark("recv_frame", msg->meta.timestamp, msg->meta.request * 2 + msg->meta.push);
        return bytes;
      }
      // sockets epoll reported readable and that we have not looked at yet; only
      // this thread reads them, so each still holds at least one byte
      while (!ready_fds_.empty()) {
        const int fd = ready_fds_.front();
        ready_fds_.pop_front();
        int bytes = ReadFrame(fd, msg);
        if (bytes > 0) return bytes;  // 0: the peer closed, rang a doorbell or offered a ring
      }
      if (int bytes = PollPipes(msg)) return bytes;
      // a derived van with asynchronous receives in flight keeps this thread polling
      // ... and so does a ring whose head frame waits for a completion only the device signals
      int timeout_ms = (HasDeferred() || gates_pending_) ? 0 : -1;
      if (timeout_ms != 0 && !pipe_fds_.empty()) {
        // stay hot for a short while after the last message, then declare ourselves asleep
        // on every ring so that the next writer rings the doorbell
        const auto idle = std::chrono::steady_clock::now() - last_activity;
        if (idle < std::chrono::microseconds(spin_window_us)) {
          timeout_ms = 0;
        } else if (!SleepOnPipes()) {
          last_activity = std::chrono::steady_clock::now();
          continue;  // something arrived while we were getting ready to sleep
        }
      }
      if (timeout_ms == 0 && (++spin_polls_ & 15) != 0) {
        // busy phase: the rings and the loopback queue are checked every pass, the sockets
        // (control traffic, doorbells, new connections) only every 16th: epoll_wait is a
        // system call even when it returns at once. Past the floor window the polling turns
        // polite (sched_yield): a thread with real work must get this CPU if it needs one
        if (std::chrono::steady_clock::now() - last_activity < std::chrono::microseconds(pipe_budget_.floor_us())) {
          CpuRelax();
        } else {
          std::this_thread::yield();
        }
        continue;
      }
      struct epoll_event evs[16];
      if (timeout_ms != 0) ++blocking_waits_;
      int n = epoll_wait(epfd_, evs, 16, timeout_ms);
      if (timeout_ms != 0) {
        WakeFromPipes();
        last_activity = std::chrono::steady_clock::now();
      }
      if (n < 0) {
        if (errno == EINTR) continue;
        LOG(WARNING) << "epoll_wait: " << strerror(errno);
        return -1;
      }
      for (int i = 0; i < n; ++i) {
        const int fd = evs[i].data.fd;
        if (fd == wake_fd_) {
          uint64_t junk;
          ssize_t r = read(wake_fd_, &junk, sizeof(junk));
          (void)r;
        } else if (fd == listen_fd_) {
          AcceptOne();
        } else if (std::find(ready_fds_.begin(), ready_fds_.end(), fd) == ready_fds_.end()) {
          ready_fds_.push_back(fd);  // level-triggered: drained one frame at a time
        }
      }
      if (n > 0) last_activity = std::chrono::steady_clock::now();
    }
  }

 private:
  socklen_t UnixAddr(int port, struct sockaddr_un* sa) {
    memset(sa, 0, sizeof(*sa));
    sa->sun_family = AF_UNIX;
    // abstract namespace: no filesystem entry to clean up
    std::string name = "pslite_b200_" + std::to_string(port);
    sa->sun_path[0] = '\0';
    memcpy(sa->sun_path + 1, name.data(), name.size());
    return static_cast<socklen_t>(offsetof(struct sockaddr_un, sun_path) + 1 + name.size());
  }
  int ConnectUnix(int port) {
    int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return -1;
    struct sockaddr_un sa;
    socklen_t len = UnixAddr(port, &sa);
    if (connect(fd, reinterpret_cast<struct sockaddr*>(&sa), len) != 0) {
      close(fd);
      return -1;
    }
    return fd;
  }
  int ConnectTcp(const std::string& host, int port) {
    struct addrinfo hints, *res = nullptr;
    memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || !res) {
      return -1;
    }
    int fd = socket(res->ai_family, res->ai_socktype | SOCK_CLOEXEC, res->ai_protocol);
    if (fd < 0) {
      freeaddrinfo(res);
      return -1;
    }
    if (connect(fd, res->ai_addr, res->ai_addrlen) != 0) {
      close(fd);
      freeaddrinfo(res);
      return -1;
    }
    freeaddrinfo(res);
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    int buf = 8 << 20;
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof(buf));
    return fd;
  }

  void AddToEpoll(int fd) {
    struct epoll_event ev;
    memset(&ev, 0, sizeof(ev));
    ev.events = EPOLLIN;
    ev.data.fd = fd;
    CHECK_EQ(epoll_ctl(epfd_, EPOLL_CTL_ADD, fd, &ev), 0) << strerror(errno);
  }

  void AcceptOne() {
    // the listen socket is blocking; accept exactly the one pending connection
    int fd = accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
    if (fd < 0) return;
    if (!local_ipc_) {
      int one = 1;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      int buf = 8 << 20;
      setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof(buf));
    }
    inbound_[fd].reset(new Inbound());
    AddToEpoll(fd);
  }

  /*! \brief create a ring for this connection and offer it to the peer over the socket */
  void OfferPipe(Peer* peer, int peer_id) {
    static std::atomic<int> seq{0};
    const std::string name = "/" + ShmScopedPrefix("pslb200_") + std::to_string(getpid()) + "_" +
                             std::to_string(peer_id) + "_" + std::to_string(seq++);
    std::unique_ptr<ShmPipe> pipe = ShmPipe::Create(name, pipe_bytes_);
    if (!pipe) return;  // no /dev/shm: stay on the socket
    FrameHeader hello = {kPipeMagic, my_node_.id, peer_id, static_cast<uint32_t>(name.size()), 0, 0, 0};
    struct iovec iov[2] = {{&hello, sizeof(hello)}, {const_cast<char*>(name.data()), name.size()}};
    if (!SendAll(peer->fd, iov, 2)) return;
    const int fd = peer->fd;
    ShmPipe* raw = pipe.get();
    // a sleeping reader never drains a full ring: wake it from inside the blocked write
    pipe->set_full_hook([fd, raw] {
      if (raw->ReaderNeedsDoorbell()) RingDoorbell(fd);
    });
    peer->offered_gate_word = MapGateWord(raw);
    peer->offered = std::move(pipe);
    peer->made_offer = true;
    {
      std::lock_guard<std::mutex> lk(offer_mu_);
      offer_fds_[fd] = peer_id;
    }
    AddToEpoll(fd);  // the answer arrives on this (otherwise write-only) socket
  }

  /*! \brief receive thread: the peer answered our offer on outbound socket `fd` */
  void OnPipeAnswer(int fd, int peer_id) {
    char answer = 0;
    const ssize_t r = recv(fd, &answer, 1, MSG_DONTWAIT);
    if (r < 0 && (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR)) return;
    epoll_ctl(epfd_, EPOLL_CTL_DEL, fd, nullptr);  // one answer per connection (or EOF)
    {
      std::lock_guard<std::mutex> lk(offer_mu_);
      offer_fds_.erase(fd);
    }
    std::shared_ptr<Peer> peer;
    {
      std::lock_guard<SpinMutex> lk(peers_mu_);
      auto it = peers_.find(peer_id);
      if (it != peers_.end()) peer = it->second;
    }
    if (!peer) return;
    if (r == 1 && answer != 'A') OnPipeVerdict(peer_id, false);  // (before the offer stops being "pending")
    std::lock_guard<std::mutex> lk(peer->mu);
    if (peer->fd != fd || !peer->offered) return;  // a reconnect replaced this connection
    if (r == 1 && answer == 'A') {
      FrameHeader sw = {kPipeSwitchMagic, my_node_.id, peer_id, 0, 0, 0, 0};
      struct iovec iov[1] = {{&sw, sizeof(sw)}};
      if (SendAll(fd, iov, 1)) {
        peer->gate_word = peer->offered_gate_word;
        peer->pipe = std::move(peer->offered);
        peer->offered_gate_word = nullptr;
        return;
      }
    } else if (r == 1) {
      PS_VLOG(1) << "node " << peer_id << " cannot map our shared-memory ring (private /dev/shm?): "
                 << "descriptors for it stay on the socket";
    }
    if (peer->offered_gate_word) ReleaseGateWord(peer->offered.get());
    peer->offered_gate_word = nullptr;
    peer->offered.reset();  // unlinks the name
  }

  static void RingDoorbell(int fd) {
    const char b = 1;
    ssize_t r = send(fd, &b, 1, MSG_NOSIGNAL);
    (void)r;
  }

  /*! \brief write whole frames into the peer's ring; doorbell only if the reader sleeps */
  bool SendThroughPipe(Peer* peer, const struct iovec* iov, int niov) {
    if (!peer->pipe->WriteV(iov, niov)) return false;
    if (peer->pipe->ReaderNeedsDoorbell()) RingDoorbell(peer->fd);
    return true;
  }

  static size_t ElemSize(DataType t) {
    switch (t) {
      case INT16: case UINT16: return 2;
      case INT32: case UINT32: case FLOAT: return 4;
      case INT64: case UINT64: case DOUBLE: return 8;
      default: return 1;
    }
  }

  static bool SendAll(int fd, struct iovec* iov, int niov) {
    while (niov > 0) {
      struct msghdr mh;
      memset(&mh, 0, sizeof(mh));
      mh.msg_iov = iov;
      mh.msg_iovlen = static_cast<size_t>(niov);
      ssize_t n = sendmsg(fd, &mh, MSG_NOSIGNAL);
      if (n < 0) {
        if (errno == EINTR) continue;
        return false;
      }
      size_t left = static_cast<size_t>(n);
      while (niov > 0 && left >= iov->iov_len) {
        left -= iov->iov_len;
        ++iov;
        --niov;
      }
      if (niov > 0 && left) {
        iov->iov_base = static_cast<char*>(iov->iov_base) + left;
        iov->iov_len -= left;
      }
    }
    return true;
  }

  /*! \brief 1 = got n bytes, 0 = orderly close before the first byte, -1 = error */
  static int ReadAll(int fd, void* dst, size_t n) {
    char* p = static_cast<char*>(dst);
    size_t got = 0;
    while (got < n) {
      ssize_t r = recv(fd, p + got, n - got, 0);
      if (r == 0) return got == 0 ? 0 : -1;
      if (r < 0) {
        if (errno == EINTR) continue;
        return -1;
      }
      got += static_cast<size_t>(r);
    }
    return 1;
  }

  // ---- outbox: sends of receive threads that would have had to wait ------------------------------
  struct Outgoing {
    std::shared_ptr<Peer> peer;
    Message msg;
  };
  void Defer(const std::shared_ptr<Peer>& peer, const Message& msg) {
    peer->deferred.fetch_add(1, std::memory_order_acq_rel);
    {
      std::lock_guard<std::mutex> lk(outbox_mu_);
      outbox_.push_back(Outgoing{peer, msg});
      if (!outbox_thread_) outbox_thread_.reset(new std::thread([this] { OutboxLoop(); }));
    }
    outbox_cv_.notify_one();
    ++deferred_sends_;
  }
  void OutboxLoop() {
    tls_outbox_ = true;  // this thread MAY wait: that is its job
    std::unique_lock<std::mutex> lk(outbox_mu_);
    for (;;) {
      outbox_cv_.wait(lk, [this] { return outbox_stop_ || !outbox_.empty(); });
      if (outbox_.empty()) return;  // (stopping, nothing left)
      Outgoing o = std::move(outbox_.front());
      outbox_.pop_front();
      lk.unlock();
      if (SendFrame(o.msg) < 0 && !outbox_stop_) {
        LOG(WARNING) << "deferred message for node " << o.msg.meta.recver << " could not be sent";
      }
      o.peer->deferred.fetch_sub(1, std::memory_order_acq_rel);
      lk.lock();
    }
  }
  void StopOutbox() {
    std::unique_ptr<std::thread> t;
    {
      std::lock_guard<std::mutex> lk(outbox_mu_);
      outbox_stop_ = true;
      t.swap(outbox_thread_);
    }
    outbox_cv_.notify_all();
    if (t) t->join();
    std::lock_guard<std::mutex> lk(outbox_mu_);
    outbox_stop_ = false;  // the van may be started again
  }

  /*! \brief deliver the next frame of any ring that has bytes; 0 if all are empty */
  int PollPipes(Message* msg) {
    const size_t n = pipe_fds_.size();
    for (size_t k = 0; k < n; ++k) {
      const size_t idx = (pipe_cursor_ + k) % n;
      auto it = inbound_.find(pipe_fds_[idx]);
      if (it == inbound_.end() || !it->second->pipe) continue;
      ShmPipe* pipe = it->second->pipe.get();
      if (pipe->Readable() == 0) continue;
      FrameHeader hdr;
      // a gated frame stays in the ring until the copy engine has signalled its payload; the
      // frames behind it wait too (a pull must not overtake the push it follows)
      if (pipe->Peek(&hdr, sizeof(hdr)) && hdr.gate != 0 && pipe->gate_done() < hdr.gate) {
        gates_pending_ = true;
        // a gate that stays closed is the face of a lost completion (a copy that was never issued, a
        // dead sender): say so instead of stalling silently (PS_WAIT_WARN_S seconds, 0 = never)
        static const int warn_s = GetEnv("PS_WAIT_WARN_S", 60);
        if (warn_s > 0) {
          Inbound* in = it->second.get();
          const auto now = std::chrono::steady_clock::now();
          if (in->gate_waiting_for != hdr.gate) {
            in->gate_waiting_for = hdr.gate;
            in->gate_since = now;
          } else if (now - in->gate_since > std::chrono::seconds(warn_s)) {
            in->gate_since = now;
            LOG(WARNING) << "node " << my_node_.id << ": the descriptor ring from node " << hdr.sender
                         << " has waited " << warn_s << " s for completion " << hdr.gate << " (reached: "
                         << pipe->gate_done() << "); everything behind it is held back";
          }
        }
        continue;
      }
      pipe_cursor_ = (idx + 1) % n;
      return ReadFramePipe(pipe, msg);
    }
    return 0;
  }
  /*! \brief flag every ring "reader asleep"; false (flags cleared) if one has data after all */
  bool SleepOnPipes() {
    EventTrace::Mark("recv_sleep", pipe_budget_.window_us());
    for (int fd : pipe_fds_) {
      auto it = inbound_.find(fd);
      if (it == inbound_.end() || !it->second->pipe) continue;
      if (!it->second->pipe->PrepareSleep()) {
        WakeFromPipes();
        return false;
      }
    }
    return true;
  }
  void WakeFromPipes() {
    for (int fd : pipe_fds_) {
      auto it = inbound_.find(fd);
      if (it != inbound_.end() && it->second->pipe) it->second->pipe->CancelSleep();
    }
  }

  void DropInbound(int fd) {
    pipe_fds_.erase(std::remove(pipe_fds_.begin(), pipe_fds_.end(), fd), pipe_fds_.end());
    epoll_ctl(epfd_, EPOLL_CTL_DEL, fd, nullptr);
    close(fd);
    inbound_.erase(fd);
  }

  SArray<char> AllocSegment(size_t len) {
    SArray<char> seg;
    if (len == 0) return seg;
    if (len >= 4096) return pool_->Get(len);
    return SArray<char>::Compact(len);  // keys / lens: one allocation, not two
  }

  /*!
   * \brief per-connection read buffer: a descriptor-sized frame (header + segment table
   *        + meta + key + len) is consumed with ONE recv() instead of five
   */
  struct Inbound {
    static constexpr size_t kCap = 64 * 1024;
    std::unique_ptr<char[]> buf{new char[kCap]};
    size_t head = 0, tail = 0;
    size_t avail() const { return tail - head; }
    /*! \brief once the peer switched to its ring, its frames arrive here and the socket only rings */
    std::unique_ptr<ShmPipe> pipe;
    std::unique_ptr<ShmPipe> offered;  // mapped and accepted, the switch marker is still to come
    uint64_t gate_waiting_for = 0;     // the gated frame at the head of the ring, and since when it waits
    std::chrono::steady_clock::time_point gate_since;
  };

  /*! \brief make >= n bytes available in the buffer (n <= kCap). 1 ok, 0 closed, -1 error */
  int Fill(int fd, Inbound* in, size_t n) {
    if (in->avail() >= n) return 1;
    if (in->head && in->head + n > Inbound::kCap) {
      memmove(in->buf.get(), in->buf.get() + in->head, in->avail());
      in->tail -= in->head;
      in->head = 0;
    } else if (in->avail() == 0) {
      in->head = in->tail = 0;
    }
    const bool first_byte = in->avail() == 0;
    while (in->avail() < n) {
      ssize_t r = recv(fd, in->buf.get() + in->tail, Inbound::kCap - in->tail, 0);
      if (r == 0) return (first_byte && in->avail() == 0) ? 0 : -1;
      if (r < 0) {
        if (errno == EINTR) continue;
        return -1;
      }
      in->tail += static_cast<size_t>(r);
    }
    return 1;
  }

  /*! \brief copy n bytes of the stream to dst: buffered bytes first, then straight from the socket */
  int Take(int fd, Inbound* in, void* dst, size_t n) {
    char* out = static_cast<char*>(dst);
    const size_t from_buf = std::min(n, in->avail());
    if (from_buf) {
      memcpy(out, in->buf.get() + in->head, from_buf);
      in->head += from_buf;
    }
    if (n == from_buf) return 1;
    if (n - from_buf <= 4096) {  // small tail: go through the buffer to batch with what follows
      int rc = Fill(fd, in, n - from_buf);
      if (rc <= 0) return -1;
      memcpy(out + from_buf, in->buf.get() + in->head, n - from_buf);
      in->head += n - from_buf;
      return 1;
    }
    return ReadAll(fd, out + from_buf, n - from_buf) == 1 ? 1 : -1;
  }

  /*! \brief read one whole frame from fd; >0 bytes, 0 if nothing to deliver (closed / doorbell) */
  int ReadFrame(int fd, Message* msg) {
    auto iit = inbound_.find(fd);
    if (iit == inbound_.end()) {
      int offered_to = -1;
      {
        std::lock_guard<std::mutex> lk(offer_mu_);
        auto oit = offer_fds_.find(fd);
        if (oit != offer_fds_.end()) offered_to = oit->second;
      }
      if (offered_to >= 0) OnPipeAnswer(fd, offered_to);
      return 0;
    }
    Inbound* in = iit->second.get();
    if (in->pipe) {
      // the socket of a pipe-backed connection only carries 1-byte doorbells (or EOF)
      char junk[256];
      ssize_t r = recv(fd, junk, sizeof(junk), MSG_DONTWAIT);
      if (r == 0) DropInbound(fd);
      return 0;  // RecvMsg polls the ring next
    }
    FrameHeader hdr;
    int rc = Fill(fd, in, sizeof(hdr));
    if (rc <= 0) {
      DropInbound(fd);
      return 0;
    }
    memcpy(&hdr, in->buf.get() + in->head, sizeof(hdr));
    in->head += sizeof(hdr);
    if (hdr.magic == kPipeMagic) {
      std::string name(hdr.meta_len, '\0');
      CHECK_EQ(Take(fd, in, &name[0], name.size()), 1);
      in->offered = ShmPipe::Attach(name);
      // PS_TEST_DECLINE_PIPE: behave like a container that cannot see the sender's /dev/shm
      static const bool decline_all = GetEnv("PS_TEST_DECLINE_PIPE", 0) != 0;
      if (decline_all) in->offered.reset();
      if (in->offered) in->offered->Unlink();  // both ends have it mapped: the name is no longer needed
      const char answer = in->offered ? 'A' : 'D';
      if (!in->offered) {
        PS_VLOG(1) << "cannot map the shared-memory ring node " << hdr.sender << " offered (" << name
                   << ": " << strerror(errno) << "): its descriptors stay on the socket";
      }
      if (!in->offered) OnPipeVerdict(hdr.sender, false);  // (before the peer's offer counts as "seen")
      {
        std::lock_guard<std::mutex> lk(offer_mu_);
        offers_seen_.insert(hdr.sender);
      }
      ssize_t w = send(fd, &answer, 1, MSG_NOSIGNAL);
      (void)w;
      if (in->avail() > 0) ready_fds_.push_back(fd);
      return 0;
    }
    if (hdr.magic == kPipeSwitchMagic) {
      // everything the peer sent before this marker has been read from the socket: from here
      // on its frames are in the ring
      CHECK(in->offered) << "ring switch without an accepted offer from node " << hdr.sender;
      in->pipe = std::move(in->offered);
      pipe_fds_.push_back(fd);
      return 0;
    }
    const int bytes = ParseFrame(hdr, msg, [&](void* dst, size_t n) { return Take(fd, in, dst, n) == 1; });
    // frames already sitting in the buffer will not wake epoll: keep this fd runnable
    if (in->avail() > 0) ready_fds_.push_back(fd);
    return bytes;
  }

  /*! \brief next frame from a ring, if one has started to arrive */
  int ReadFramePipe(ShmPipe* pipe, Message* msg) {
    FrameHeader hdr;
    CHECK(pipe->Read(&hdr, sizeof(hdr))) << "shared-memory ring writer vanished";
    const int rc = ParseFrame(hdr, msg, [&](void* dst, size_t n) { return pipe->Read(dst, n); });
    pipe->Commit();  // hand the frame's space back with one store
    return rc;
  }

  /*!
   * \brief where segment `i` (len bytes) of an arriving message goes: the buffer registered for
   *        (sender, key), the pull's own destination (zero-copy pull), else fresh memory
   */
  SArray<char> SegmentDestination(const Meta& meta, uint32_t i, uint64_t len) {
    SArray<char> seg;
    if (i == 1 && meta.push && meta.request) {
      std::lock_guard<std::mutex> lk(reg_mu_);
      auto it = registered_.find(std::make_pair(meta.sender, meta.key));
      if (it != registered_.end()) {
        CHECK_GE(it->second.size(), len) << "registered buffer too small";
        seg = it->second.segment(0, len);
      }
    }
    // a pull response can land straight in the buffer the request named (meta.addr
    // is this process's own pointer, echoed back by the server): zero-copy pull
    // — but only the buffer THIS van recorded when it sent the request: a frame from the network
    // must never be able to name an address to write to, and a late duplicate of a reply that
    // was already consumed finds no record and lands in fresh memory
    if (i == 1 && direct_pull_ && !meta.request && !meta.push && meta.addr != 0 && len > 0 &&
        meta.src_dev_type != GPU) {
      PullDest rec = {0, 0};
      {
        std::lock_guard<SpinMutex> lk(pull_mu_);
        auto it = pull_dests_.find(PullKey(meta.sender, meta.app_id, meta.customer_id, meta.timestamp));
        if (it != pull_dests_.end()) {
          rec = it->second;
          pull_dests_.erase(it);
        }
      }
      if (rec.addr != 0 && rec.addr == meta.addr && len <= rec.bytes) {
        seg.reset(reinterpret_cast<char*>(rec.addr), len, [](char*) {});
      }
    }
    if (seg.size() != len) seg = AllocSegment(len);
    // the bytes (will) live in host memory of this node
    seg.src_device_type_ = CPU;
    seg.src_device_id_ = 0;
    seg.dst_device_type_ = i == 1 && meta.dst_dev_type != UNK ? meta.dst_dev_type : CPU;
    seg.dst_device_id_ = i == 1 && meta.dst_dev_id >= 0 ? meta.dst_dev_id : 0;
    return seg;
  }

  /*!
   * \brief in-process hand-off (sender side, any thread): build the message the peer van would
   *        have parsed from our frame — segments copied into ITS destinations — and give it to
   *        that van directly: no serialisation, no ring, no hop through its receive thread.
   */
  bool HandOff(Peer* peer, const Message& msg) {
    std::shared_lock<std::shared_mutex> lk(LocalVans().mu);
    auto it = LocalVans().by_port.find(peer->port);
    if (it == LocalVans().by_port.end()) return false;
    TcpVan* dst = it->second;
    if (dst == this || !dst->AllowLocalHandoff()) return false;
    Message m;
    m.meta = msg.meta;
    m.meta.sender = my_node_.id;
    m.meta.recver = dst->my_node_.id;
    if (m.meta.recver != msg.meta.recver) return false;  // stale registry entry
    for (uint32_t i = 0; i < msg.data.size(); ++i) {
      const SArray<char>& src = msg.data[i];
      CHECK(src.size() == 0 || !src.on_gpu()) << "TcpVan cannot send device memory";
      SArray<char> seg = dst->SegmentDestination(m.meta, i, src.size());
      if (src.size() && seg.data() != src.data()) memcpy(seg.data(), src.data(), src.size());
      m.data.push_back(seg);
    }
    if (!m.meta.control.empty()) {
      // rendezvous traffic of a derived transport: it must stay ordered with the data
      // messages it prepares (a pull request must not overtake its region announcement)
      if (!dst->OnLocalControl(&m)) return false;
      ++handoffs_;
      return true;
    }
    dst->OnLocalDeliver(&m);
    if (!dst->AcceptHandoff(&m)) return false;
    ++handoffs_;
    return true;
  }

  /*! \brief everything after the FrameHeader; `take(dst, n)` pulls the next n stream bytes */
  template <typename TakeFn>
  int ParseFrame(const FrameHeader& hdr, Message* msg, TakeFn take) {
    CHECK_EQ(hdr.magic, kFrameMagic) << "corrupt frame";
    CHECK_LE(hdr.num_segments, kMaxSegments);
    uint64_t seg_len[kMaxSegments];
    if (hdr.num_segments) CHECK(take(seg_len, sizeof(uint64_t) * hdr.num_segments));
    thread_local std::vector<char> meta_buf;  // reused: no allocation per frame
    meta_buf.resize(hdr.meta_len);
    CHECK(take(meta_buf.data(), meta_buf.size()));
    CHECK(UnpackMeta(meta_buf.data(), meta_buf.size(), &msg->meta)) << "corrupt meta";
    msg->meta.sender = hdr.sender;
    msg->meta.recver = my_node_.id;
    size_t total = sizeof(hdr) + meta_buf.size();
    msg->data.clear();
    msg->data.reserve(hdr.num_segments);
    for (uint32_t i = 0; i < hdr.num_segments; ++i) {
      SArray<char> seg = SegmentDestination(msg->meta, i, seg_len[i]);
      if (seg_len[i]) CHECK(take(seg.data(), seg_len[i]));
      msg->data.push_back(seg);
      total += seg_len[i];
    }
    return static_cast<int>(std::min<size_t>(total, 0x7fffffff));
  }

 protected:
  /*!
   * \brief hooks for transports whose payload arrives asynchronously (NcclVan): called on the
   *        receive thread at the top of every RecvMsg iteration; return true with a message
   *        whose payload has completed. HasDeferred() == true turns the idle wait into a poll.
   */
  /*! \brief may this van take part in in-process hand-offs (both ends must agree)? */
  virtual bool AllowLocalHandoff() const { return true; }
  /*! \brief what the derived RecvMsg would have done to a message that was handed off */
  virtual void OnLocalDeliver(Message* /*msg*/) {}
  /*! \brief handle a transport-level control message (ADDR_*) handed off in-process; false = not mine */
  virtual bool OnLocalControl(Message* /*msg*/) { return false; }
  virtual bool PollDeferred(Message* /*msg*/) { return false; }
  virtual bool HasDeferred() { return false; }
  /*! \brief a ring to a same-host peer was created: return where the copy engine signals the
   *  completion of gated frames (see ShmPipe::gate_word), or null for "no gating" */
  virtual void* MapGateWord(ShmPipe* /*pipe*/) { return nullptr; }
  /*! \brief the ring of a gated connection is about to be unmapped */
  virtual void ReleaseGateWord(ShmPipe* /*pipe*/) {}
  /*! \brief enqueue a message for this van's own RecvMsg and wake it */
  int Loopback(const Message& msg) {
    {
      std::lock_guard<std::mutex> lk(loop_mu_);
      loop_q_.push_back(msg);
      loop_q_.back().meta.sender = my_node_.id;
    }
    uint64_t one = 1;
    ssize_t r = write(wake_fd_, &one, sizeof(one));
    (void)r;
    return 1 + static_cast<int>(msg.meta.data_size & 0x3fffffff);
  }
 private:
  bool PopLoopback(Message* msg) {
    std::lock_guard<std::mutex> lk(loop_mu_);
    if (loop_q_.empty()) return false;
    *msg = std::move(loop_q_.front());
    loop_q_.pop_front();
    return true;
  }

  void CloseAll() {
    std::lock_guard<std::mutex> ilk(init_mu_);
    {
      std::lock_guard<SpinMutex> lk(peers_mu_);
      for (auto& kv : peers_) {
        std::lock_guard<std::mutex> plk(kv.second->mu);
        if (kv.second->fd >= 0) close(kv.second->fd);
        kv.second->fd = -1;
        if (kv.second->pipe && kv.second->gate_word) {
          ReleaseGateWord(kv.second->pipe.get());
          kv.second->gate_word = nullptr;
          kv.second->gate_keep.clear();
        }
        if (kv.second->offered && kv.second->offered_gate_word) {
          ReleaseGateWord(kv.second->offered.get());
          kv.second->offered_gate_word = nullptr;
        }
      }
      peers_.clear();
    }
    for (auto& kv : inbound_) close(kv.first);
    inbound_.clear();
    pipe_fds_.clear();
    ready_fds_.clear();
    if (listen_fd_ >= 0) close(listen_fd_);
    listen_fd_ = -1;
    if (wake_fd_ >= 0) close(wake_fd_);
    wake_fd_ = -1;
    if (epfd_ >= 0) close(epfd_);
    epfd_ = -1;
    {
      std::lock_guard<std::mutex> lk(loop_mu_);
      loop_q_.clear();
    }
  }

  std::mutex init_mu_;
  std::shared_ptr<RecvBufferPool> pool_;
  bool direct_pull_ = true;
  bool local_ipc_ = false;
  int connect_timeout_s_ = 120;
  int epfd_ = -1;
  int wake_fd_ = -1;
  int listen_fd_ = -1;
  std::unordered_map<int, std::unique_ptr<Inbound>> inbound_;  // receive thread only
  static void CpuRelax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  /*! \brief the TcpVans of this process, by listening port */
  struct LocalVanTable {
    std::shared_mutex mu;
    std::map<int, TcpVan*> by_port;
  };
  static LocalVanTable& LocalVans() {
    static LocalVanTable* t = new LocalVanTable();
    return *t;
  }
  /*!
   * \brief PS_LOCAL_HANDOFF=1: data messages between vans of one process skip the wire format and
   *        the receiver's van thread. Off by default: it removes serialisation (-14 % per key on
   *        the TCP van, joint 1 KB) but moves the receive work onto the sending thread, which
   *        costs more than it saves once that thread also issues the copies (+10 % on the
   *        one-sided shm van) — measured in the development container.
   */
  bool handoff_ = GetEnv("PS_LOCAL_HANDOFF", 0) != 0;
  std::atomic<uint64_t> handoffs_{0};
  uint32_t spin_polls_ = 0;
  std::atomic<uint64_t> blocking_waits_{0};                    // times the receive thread went to sleep in epoll
  bool gates_pending_ = false;                                 // receive thread: a ring head waits for its gate
  std::vector<int> pipe_fds_;                                  // inbound connections with a ring
  size_t pipe_cursor_ = 0;
  bool use_pipes_ = true;
  size_t pipe_bytes_ = 256u << 10;
  SpinBudget pipe_budget_{GetEnv("PS_SHM_PIPE_SPIN_US", 50), GetEnv("PS_SPIN_MAX_US", 1000)};
  std::deque<int> ready_fds_;      // touched by the receive thread only
  SpinMutex peers_mu_;  // a map lookup per send
  std::unordered_map<int, std::shared_ptr<Peer>> peers_;
  std::mutex loop_mu_;
  std::deque<Message> loop_q_;
  std::mutex reg_mu_;
  std::map<std::pair<int, uint64_t>, SArray<char>> registered_;
  /*! \brief destinations of the pulls in flight, by (server, app, customer, timestamp) */
  struct PullDest {
    uint64_t addr;
    uint64_t bytes;
  };
  using PullKey = std::tuple<int, int, int, int>;
  static constexpr size_t kMaxPullRecords = 1 << 16;
  std::mutex outbox_mu_;
  std::condition_variable outbox_cv_;
  std::deque<Outgoing> outbox_;
  std::unique_ptr<std::thread> outbox_thread_;
  bool outbox_stop_ = false;
  std::atomic<uint64_t> deferred_sends_{0};
  static thread_local const TcpVan* tls_receiving_;  // set on a van's receive thread
  static thread_local bool tls_outbox_;               // set on an outbox thread
  std::mutex offer_mu_;
  std::unordered_map<int, int> offer_fds_;  // outbound socket -> peer id, while an offer is unanswered
  std::set<int> offers_seen_;               // peers whose own ring offer has arrived (under offer_mu_)
  SpinMutex pull_mu_;
  std::map<PullKey, PullDest> pull_dests_;
};

inline thread_local const TcpVan* TcpVan::tls_receiving_ = nullptr;
inline thread_local bool TcpVan::tls_outbox_ = false;

}  // namespace ps

#endif  // PS_VAN_TCP_VAN_H_
