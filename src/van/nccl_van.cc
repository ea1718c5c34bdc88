/**
 * \file nccl_van.cc
 * \brief NcclVan implementation (see nccl_van.h for the protocol and the reference mapping).
 *
 * libnccl is resolved with dlopen at van start — the library, the torch extension and the CPU
 * test binaries keep no link-time dependency on it — and only the seven entry points below
 * are used. Threading: application threads send (one critical section per directed link keeps
 * descriptor order == ncclSend order); the van's receive thread posts every ncclRecv and is
 * also the only thread that delivers, through TcpVan's PollDeferred / HasDeferred hooks, so
 * the deferred queue needs no lock.
 */
#include "van/nccl_van.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "kernels/ps_kernels.h"
#include "ps/internal/postoffice.h"
#include "van/cuda_domain.h"
#include "van/mem_domain.h"
#include "van/tcp_van.h"

namespace ps {
namespace {

#define NV_CUDA(expr)                                                             \
  do {                                                                            \
    cudaError_t e_ = (expr);                                                      \
    CHECK(e_ == cudaSuccess) << #expr << ": " << cudaGetErrorString(e_);          \
  } while (0)

/*! \brief MemRef::region markers of this van (kSymmetricRegion is 0x40000000) */
const int32_t kViaNccl = 0x40000001;    // the payload follows on the pair's communicator
const int32_t kViaHost = 0x40000002;    // the payload is in the frame (host staged); place it on the GPU

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;

  /*! \brief process-wide table; nullptr if no usable libnccl can be loaded */
  static const NcclApi* Get() {
    static const NcclApi* api = Load();
    return api;
  }

 private:
  static const NcclApi* Load() {
    // a library already mapped under this soname (PyTorch bundles one) is reused by dlopen
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      LOG(ERROR) << "the nccl van needs libnccl.so.2: " << dlerror();
      return nullptr;
    }
    std::unique_ptr<NcclApi> a(new NcclApi());
    bool ok = true;
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) {
        LOG(ERROR) << "libnccl lacks " << n;
        ok = false;
      }
      return p;
    };
    a->GetUniqueId = reinterpret_cast<decltype(a->GetUniqueId)>(sym("ncclGetUniqueId"));
    a->CommInitRank = reinterpret_cast<decltype(a->CommInitRank)>(sym("ncclCommInitRank"));
    a->CommAbort = reinterpret_cast<decltype(a->CommAbort)>(sym("ncclCommAbort"));
    a->Send = reinterpret_cast<decltype(a->Send)>(sym("ncclSend"));
    a->Recv = reinterpret_cast<decltype(a->Recv)>(sym("ncclRecv"));
    a->GetErrorString = reinterpret_cast<decltype(a->GetErrorString)>(sym("ncclGetErrorString"));
    a->GetVersion = reinterpret_cast<decltype(a->GetVersion)>(sym("ncclGetVersion"));
    return ok ? a.release() : nullptr;
  }
};

#define NV_NCCL(expr)                                                                   \
  do {                                                                                  \
    ncclResult_t r_ = (expr);                                                           \
    CHECK(r_ == ncclSuccess) << #expr << ": " << NcclApi::Get()->GetErrorString(r_);   \
  } while (0)

/*! \brief cudaEvent that destroys itself with its last owner */
struct EventBox {
  cudaEvent_t ev = nullptr;
  EventBox() { NV_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
  ~EventBox() {
    if (ev) cudaEventDestroy(ev);
  }
  bool Done() const {
    cudaError_t e = cudaEventQuery(ev);
    if (e == cudaErrorNotReady) {
      cudaGetLastError();
      return false;
    }
    return true;  // finished — or failed, which the next CUDA call reports
  }
};

class NcclVan : public TcpVan {
 public:
  NcclVan(Postoffice* postoffice, MemDomain* domain)
      : TcpVan(postoffice), api_(NcclApi::Get()), domain_(domain), dev_(domain->device()) {}
  ~NcclVan() override { ReleaseLinks(); }

  std::string GetType() const override { return "nccl"; }

  struct PeerInfo {
    int dev = -1;            // CUDA ordinal the peer drives, -1 = none
    bool same_host = false;
  };

  void Start(int customer_id, bool standalone) override {
    NV_CUDA(cudaSetDevice(dev_));
    if (!host_stream_) NV_CUDA(cudaStreamCreateWithFlags(&host_stream_, cudaStreamNonBlocking));
    int v = 0;
    if (api_->GetVersion(&v) == ncclSuccess) PS_VLOG(1) << "nccl van on GPU " << dev_ << ", NCCL " << v;
    TcpVan::Start(customer_id, standalone);
  }

  void Stop() override {
    TcpVan::Stop();  // joins the receive thread: nobody touches links / deferred_ afterwards
    ReleaseLinks();
    deferred_.clear();
    std::lock_guard<std::mutex> lk(mu_);
    landing_.clear();
    gpu_registered_.clear();
  }

  void SetNode(const Node& node) override {
    Node n = node;
    n.dev_id = dev_;
    Van::SetNode(n);
  }

  void Connect(const Node& node) override {
    {
      std::lock_guard<std::mutex> lk(mu_);
      PeerInfo& p = peers_info_[node.id];
      p.dev = node.dev_id;
      p.same_host = node.hostname == my_node_.hostname;
    }
    TcpVan::Connect(node);
  }

  void RegisterRecvBuffer(Message& msg) override {
    CHECK_GE(msg.data.size(), (size_t)2);
    if (msg.data[1].src_device_type_ != GPU) {
      TcpVan::RegisterRecvBuffer(msg);
      return;
    }
    std::lock_guard<std::mutex> lk(mu_);
    gpu_registered_[std::make_pair(msg.meta.sender, msg.meta.key)] = msg.data[1];
  }

  void* AllocExportable(size_t bytes) override { return domain_->Alloc(bytes); }
  void FreeExportable(void* p) override { domain_->Free(p); }
  /*! \brief applications enqueue their kernels here; every send is ordered behind it */
  void* DataStream() override { return domain_->Stream(); }

 protected:
  // ---------------------------------------------------------------- send side
  int SendMsg(Message& msg) override {
    if (!msg.meta.control.empty() || msg.meta.simple_app) return TcpVan::SendMsg(msg);
    const bool device_vals =
        msg.data.size() >= 2 && msg.data[1].size() > 0 && msg.data[1].src_device_type_ == GPU;
    if (!device_vals) {
      // a reply echoes the request's MemRef; it must not read as "payload follows"
      msg.meta.mem = MemRef();
      return TcpVan::SendMsg(msg);
    }
    ReapSends();
    PeerInfo peer;
    if (msg.meta.recver == my_node_.id) {
      peer.dev = dev_;
      peer.same_host = true;
    } else {
      std::lock_guard<std::mutex> lk(mu_);
      peer = peers_info_[msg.meta.recver];
    }
    const bool shares_my_gpu = peer.same_host && peer.dev == dev_;
    if (peer.dev < 0 || shares_my_gpu || msg.meta.recver == my_node_.id) return SendViaHost(msg, peer);
    return SendViaNccl(msg);
  }

  /*! \brief wire form of the values on `stream`: the tensor itself, or an encoded staging copy */
  const void* Encode(const Message& msg, uint64_t wire, cudaStream_t stream, void** stage) {
    const SArray<char>& vals = msg.data[1];
    *stage = nullptr;
    // order behind the application's kernels on DataStream() and behind the producer event
    EventBox after_app;
    NV_CUDA(cudaEventRecord(after_app.ev, static_cast<cudaStream_t>(domain_->Stream())));
    NV_CUDA(cudaStreamWaitEvent(stream, after_app.ev, 0));
    if (msg.wait_event) NV_CUDA(cudaStreamWaitEvent(stream, static_cast<cudaEvent_t>(msg.wait_event), 0));
    if (msg.meta.codec == kCodecRaw) return vals.data();
    NV_CUDA(cudaMallocAsync(stage, wire, stream));
    const int rc = ps_launch_copy(*stage, vals.data(), vals.size(), msg.meta.codec, msg.meta.scale, 0,
                                  reinterpret_cast<ps_stream_t>(stream));
    CHECK_EQ(rc, 0) << "codec kernel: " << cudaGetErrorString(static_cast<cudaError_t>(rc));
    return *stage;
  }

  static Message DescriptorOf(const Message& msg, int32_t region, uint64_t wire) {
    Message desc;
    desc.meta = msg.meta;  // keeps src_dev_type = GPU: the frame must not be landed by TcpVan
    desc.meta.mem = MemRef();
    desc.meta.mem.region = region;
    desc.meta.mem.bytes = wire;
    desc.data = msg.data;
    desc.data[1] = msg.data[1].segment(0, 0);
    desc.meta.data_size = msg.meta.data_size - static_cast<int64_t>(msg.data[1].size());
    return desc;
  }

  int SendViaNccl(Message& msg) {
    const SArray<char>& vals = msg.data[1];
    const uint64_t wire = WireBytes(msg.meta.codec, vals.size());
    Link* link = OutLink(msg.meta.recver);
    std::lock_guard<std::mutex> lk(link->mu);
    NV_CUDA(cudaSetDevice(dev_));
    EnsureOutLink(link, msg.meta.recver);
    void* stage = nullptr;
    const void* src = Encode(msg, wire, link->stream, &stage);
    Message desc = DescriptorOf(msg, kViaNccl, wire);
    // descriptor first, then the matching send, inside one critical section per link:
    // the peer posts its receives in descriptor order
    if (TcpVan::SendMsg(desc) < 0) {
      if (stage) cudaFreeAsync(stage, link->stream);
      return -1;
    }
    NV_NCCL(api_->Send(src, wire, ncclChar, 1, link->comm, link->stream));
    if (stage) NV_CUDA(cudaFreeAsync(stage, link->stream));
    InFlightSend s;
    s.done.reset(new EventBox());
    NV_CUDA(cudaEventRecord(s.done->ev, link->stream));
    s.keep_alive = vals;
    {
      std::lock_guard<std::mutex> slk(sends_mu_);
      sends_.push_back(std::move(s));
    }
    ++nccl_sends_;
    return 1 + static_cast<int>(std::min<uint64_t>(wire, 0x3fffffff));
  }

  /*! \brief peers NCCL cannot reach (same GPU, no GPU): device -> host -> frame */
  int SendViaHost(Message& msg, const PeerInfo& peer) {
    const SArray<char>& vals = msg.data[1];
    const uint64_t wire = WireBytes(msg.meta.codec, vals.size());
    SArray<char> host(wire);
    {
      std::lock_guard<std::mutex> lk(host_mu_);
      NV_CUDA(cudaSetDevice(dev_));
      void* stage = nullptr;
      const void* src = Encode(msg, wire, host_stream_, &stage);
      NV_CUDA(cudaMemcpyAsync(host.data(), src, wire, cudaMemcpyDeviceToHost, host_stream_));
      if (stage) NV_CUDA(cudaFreeAsync(stage, host_stream_));
      NV_CUDA(cudaStreamSynchronize(host_stream_));
    }
    Message desc = DescriptorOf(msg, peer.dev >= 0 ? kViaHost : -1, wire);
    host.src_device_type_ = CPU;  // what TcpVan sees; meta keeps saying GPU for the receiver
    desc.data[1] = host;
    desc.meta.data_size += static_cast<int64_t>(wire);
    if (peer.dev < 0) {  // a GPU-less receiver takes the bytes as an ordinary host payload
      desc.meta.src_dev_type = CPU;
      desc.meta.src_dev_id = 0;
    }
    ++host_sends_;
    return TcpVan::SendMsg(desc);
  }

  // ---------------------------------------------------------------- receive side
  int RecvMsg(Message* msg) override {
    for (;;) {
      delivered_from_queue_ = false;
      const int n = TcpVan::RecvMsg(msg);
      if (n < 0 || delivered_from_queue_) return n;
      if (msg->meta.control.cmd == Control::ADDR_REQUEST) {
        JoinLink(*msg);
        continue;
      }
      if (!msg->meta.control.empty()) return n;
      const int32_t how = msg->meta.mem.region;
      if (how == kViaNccl) {
        PostRecv(msg);
        continue;
      }
      if (how == kViaHost) {
        PlaceFromHost(msg);
        continue;
      }
      // ordinary message: it must not overtake payload messages that are still landing
      if (!deferred_.empty()) {
        Deferred d;
        d.msg = *msg;
        deferred_.push_back(std::move(d));
        continue;
      }
      return n;
    }
  }

  // payload messages complete asynchronously on this van's receive thread: no in-process shortcut
  bool AllowLocalHandoff() const override { return false; }

  bool HasDeferred() override { return !deferred_.empty(); }

  bool PollDeferred(Message* msg) override {
    if (deferred_.empty()) return false;
    Deferred& d = deferred_.front();
    if (d.landed && !d.landed->Done()) return false;
    *msg = std::move(d.msg);
    deferred_.pop_front();
    delivered_from_queue_ = true;
    return true;
  }

 private:
  struct Link {
    std::mutex mu;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
  };
  struct InFlightSend {
    std::unique_ptr<EventBox> done;
    SArray<char> keep_alive;
  };
  struct Deferred {
    Message msg;
    std::unique_ptr<EventBox> landed;  // null: nothing to wait for, only ordering
    SArray<char> host_keep_alive;
  };
  using PeerKey = std::pair<int, uint64_t>;

  Link* OutLink(int peer) {
    std::lock_guard<std::mutex> lk(mu_);
    std::unique_ptr<Link>& l = out_links_[peer];
    if (!l) l.reset(new Link());
    return l.get();
  }

  /*! \brief sender half of the bootstrap (link->mu held): rank 0 of a fresh 2-rank communicator */
  void EnsureOutLink(Link* link, int peer) {
    if (link->comm) return;
    ncclUniqueId id;
    NV_NCCL(api_->GetUniqueId(&id));
    Message hello;
    hello.meta.recver = peer;
    hello.meta.request = true;
    hello.meta.control.cmd = Control::ADDR_REQUEST;
    hello.meta.body.assign(id.internal, sizeof(id.internal));
    hello.meta.timestamp = GetTimestamp();
    CHECK_GT(TcpVan::SendMsg(hello), 0);
    NV_NCCL(api_->CommInitRank(&link->comm, 2, id, 0));
    NV_CUDA(cudaStreamCreateWithFlags(&link->stream, cudaStreamNonBlocking));
    PS_VLOG(1) << "nccl link " << my_node_.id << " -> " << peer << " up";
  }

  /*! \brief receiver half (van thread): rank 1; the id arrived ahead of the first descriptor */
  void JoinLink(const Message& hello) {
    CHECK_EQ(hello.meta.body.size(), sizeof(ncclUniqueId)) << "malformed nccl bootstrap message";
    ncclUniqueId id;
    memcpy(id.internal, hello.meta.body.data(), sizeof(id.internal));
    NV_CUDA(cudaSetDevice(dev_));
    std::unique_ptr<Link> link(new Link());
    NV_NCCL(api_->CommInitRank(&link->comm, 2, id, 1));
    NV_CUDA(cudaStreamCreateWithFlags(&link->stream, cudaStreamNonBlocking));
    std::unique_ptr<Link>& slot = in_links_[hello.meta.sender];
    if (slot) AbortLink(slot.get());  // the peer restarted (recovery): retire the old communicator
    slot = std::move(link);
  }

  /*! \brief bytes per element of the value segment (Meta::val_len counts elements) */
  static uint64_t ValueSize(const Message& msg) {
    if (msg.meta.data_type.size() < 2) return 1;
    switch (msg.meta.data_type[1]) {
      case INT16: case UINT16: return 2;
      case INT32: case UINT32: case FLOAT: return 4;
      case INT64: case UINT64: case DOUBLE: return 8;
      default: return 1;
    }
  }

  static bool IsDevicePointer(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
  }

  /*!
   * \brief where `wire` bytes of this message belong: the buffer registered for (sender, key),
   *        the pull's own destination tensor (zero-copy pull), or a cached landing buffer.
   *        `*host_dst` is set when a pull names a host buffer: land on the device, then copy.
   */
  char* Destination(const Message& msg, uint64_t wire, char** host_dst) {
    *host_dst = nullptr;
    const PeerKey pk(msg.meta.sender, msg.meta.key);
    if (msg.meta.request && msg.meta.push) {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = gpu_registered_.find(pk);
      if (it != gpu_registered_.end() && it->second.size() >= wire) return it->second.data();
    } else if (!msg.meta.request && !msg.meta.push && msg.meta.addr != 0 &&
               static_cast<uint64_t>(msg.meta.val_len) * ValueSize(msg) >= wire) {
      char* dst = reinterpret_cast<char*>(msg.meta.addr);
      if (IsDevicePointer(dst)) return dst;
      *host_dst = dst;
    }
    std::lock_guard<std::mutex> lk(mu_);
    std::pair<char*, uint64_t>& slot = landing_[pk];
    if (slot.second < wire) {
      // grow only; an outgrown buffer stays in the arena until Stop (Free would have to drain
      // the device, and kernels of an unfinished round may still read it)
      slot.second = AlignUp(wire, 512);
      slot.first = static_cast<char*>(domain_->Alloc(slot.second));
      CHECK(slot.first) << "out of device memory for a " << slot.second << " B landing buffer";
    }
    return slot.first;
  }

  void FinishLanding(Message* msg, char* dst, char* host_dst, uint64_t wire, cudaStream_t stream,
                     Deferred* d) {
    if (host_dst) {
      NV_CUDA(cudaMemcpyAsync(host_dst, dst, wire, cudaMemcpyDeviceToHost, stream));
      dst = host_dst;
    }
    d->landed.reset(new EventBox());
    NV_CUDA(cudaEventRecord(d->landed->ev, stream));
    SArray<char> vals;
    const bool on_dev = host_dst == nullptr;
    vals.reset(dst, wire, [](char*) {}, on_dev ? GPU : CPU, on_dev ? dev_ : 0, on_dev ? GPU : CPU,
               on_dev ? dev_ : 0);
    msg->data[1] = vals;
    msg->meta.data_size += static_cast<int64_t>(wire);
    msg->meta.mem = MemRef();
    d->msg = *msg;
  }

  void PostRecv(Message* msg) {
    CHECK_GE(msg->data.size(), (size_t)2);
    const uint64_t wire = msg->meta.mem.bytes;
    auto it = in_links_.find(msg->meta.sender);
    CHECK(it != in_links_.end()) << "nccl payload from node " << msg->meta.sender << " before its link";
    Link* link = it->second.get();
    NV_CUDA(cudaSetDevice(dev_));
    char* host_dst = nullptr;
    char* dst = Destination(*msg, wire, &host_dst);
    NV_NCCL(api_->Recv(dst, wire, ncclChar, 0, link->comm, link->stream));
    Deferred d;
    FinishLanding(msg, dst, host_dst, wire, link->stream, &d);
    deferred_.push_back(std::move(d));
    ++nccl_recvs_;
  }

  void PlaceFromHost(Message* msg) {
    CHECK_GE(msg->data.size(), (size_t)2);
    const SArray<char> host = msg->data[1];
    const uint64_t wire = host.size();
    NV_CUDA(cudaSetDevice(dev_));
    char* host_dst = nullptr;
    char* dst = Destination(*msg, wire, &host_dst);
    Deferred d;
    if (host_dst) {  // host to host: nothing for the GPU to do
      if (host_dst != host.data()) memcpy(host_dst, host.data(), wire);
      msg->meta.mem = MemRef();
      SArray<char> vals;
      vals.reset(host_dst, wire, [](char*) {});
      msg->data[1] = vals;
      d.msg = *msg;
    } else {
      msg->meta.data_size -= static_cast<int64_t>(wire);  // FinishLanding adds it back
      NV_CUDA(cudaMemcpyAsync(dst, host.data(), wire, cudaMemcpyHostToDevice, host_stream_));
      FinishLanding(msg, dst, nullptr, wire, host_stream_, &d);
      d.host_keep_alive = host;
    }
    deferred_.push_back(std::move(d));
  }

  void ReapSends() {
    std::lock_guard<std::mutex> lk(sends_mu_);
    while (!sends_.empty() && sends_.front().done->Done()) sends_.pop_front();
  }

  void AbortLink(Link* l) {
    if (l->stream) {
      cudaStreamSynchronize(l->stream);
      cudaStreamDestroy(l->stream);
      l->stream = nullptr;
    }
    if (l->comm) {
      api_->CommAbort(l->comm);  // never blocks on a peer that is already gone
      l->comm = nullptr;
    }
  }

  void ReleaseLinks() {
    cudaSetDevice(dev_);
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : out_links_) AbortLink(kv.second.get());
    for (auto& kv : in_links_) AbortLink(kv.second.get());
    out_links_.clear();
    in_links_.clear();
    {
      std::lock_guard<std::mutex> slk(sends_mu_);
      sends_.clear();
    }
    if (host_stream_) {
      cudaStreamDestroy(host_stream_);
      host_stream_ = nullptr;
    }
  }

  const NcclApi* api_;
  std::unique_ptr<MemDomain> domain_;  // device selection, landing arena, application stream
  int dev_;
  cudaStream_t host_stream_ = nullptr;
  std::mutex host_mu_;

  std::mutex mu_;
  std::map<int, PeerInfo> peers_info_;
  std::map<int, std::unique_ptr<Link>> out_links_;            // guarded by mu_ (the map), Link::mu
  std::map<int, std::unique_ptr<Link>> in_links_;             // receive thread only
  std::map<PeerKey, std::pair<char*, uint64_t>> landing_;
  std::map<PeerKey, SArray<char>> gpu_registered_;

  std::mutex sends_mu_;
  std::deque<InFlightSend> sends_;

  std::deque<Deferred> deferred_;  // receive thread only
  bool delivered_from_queue_ = false;

  std::atomic<uint64_t> nccl_sends_{0}, nccl_recvs_{0}, host_sends_{0};
};

}  // namespace

Van* CreateNcclVan(Postoffice* postoffice) {
  if (!NcclApi::Get()) return nullptr;
  MemDomain* dom = CreateCudaDomain(postoffice ? postoffice->instance_idx() : 0);
  if (!dom) return nullptr;
  return new NcclVan(postoffice, dom);
}

}  // namespace ps
