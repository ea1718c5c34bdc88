/**
 * \file gpu_server.cc
 * \brief GpuServer implementation (see gpu_server.h).
 */
#include "server/gpu_server.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>

#include "van/mem_domain.h"

namespace ps {

#define GS_CUDA(expr)                                                             \
  do {                                                                            \
    cudaError_t e_ = (expr);                                                      \
    CHECK(e_ == cudaSuccess) << "CUDA: " #expr " -> " << cudaGetErrorString(e_);  \
  } while (0)

namespace {
SArray<Key> OneKey(Key k) {
  SArray<Key> a(1);
  a[0] = k;
  return a;
}
SArray<int> OneLen(size_t bytes) {
  SArray<int> a(1);
  a[0] = static_cast<int>(std::min<size_t>(bytes, 0x7fffffff));
  return a;
}
}  // namespace

GpuServer::GpuServer(int app_id, const GpuServerConfig& cfg, int instance_idx)
    : cfg_(cfg), instance_idx_(instance_idx) {
  CHECK_GE(cfg_.num_workers, 1);
  CHECK_LE(cfg_.num_workers, PS_MAX_FANIN);
  po_ = Postoffice::GetServer(instance_idx);
  dev_ = po_->van()->my_node().dev_id;
  if (dev_ < 0) GS_CUDA(cudaGetDevice(&dev_));
  GS_CUDA(cudaSetDevice(dev_));
  stream_ = po_->van()->DataStream();
  if (!stream_) {
    cudaStream_t s;
    GS_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    stream_ = s;
  }
  server_.reset(new KVServer<char>(app_id, false, instance_idx));
  using namespace std::placeholders;
  server_->set_request_handle(std::bind(&GpuServer::Handle, this, _1, _2, _3));
}

GpuServer::~GpuServer() {
  server_.reset();
  cudaSetDevice(dev_);
  cudaStreamSynchronize(static_cast<cudaStream_t>(stream_));
  for (auto& kv : shards_) {
    cudaFree(kv.second.master);
    cudaFree(kv.second.m);
    cudaFree(kv.second.v);
  }
}

void GpuServer::SetLearningRate(float lr) {
  std::lock_guard<std::mutex> lk(mu_);
  cfg_.opt.lr = lr;
}

void GpuServer::SetSymmetricParams(void* mc_base, const std::vector<void*>& peer_bases,
                                   size_t bytes) {
  std::lock_guard<std::mutex> lk(mu_);
  CHECK_EQ(peer_bases.size(), static_cast<size_t>(cfg_.num_workers));
  mc_base_ = mc_base;
  peer_bases_ = peer_bases;
  symm_bytes_ = bytes;
}

void GpuServer::SetSymmetricGrads(void* mc_base, size_t bytes) {
  std::lock_guard<std::mutex> lk(mu_);
  mc_grad_base_ = mc_base;
  symm_grad_bytes_ = bytes;
}

size_t GpuServer::num_keys() {
  std::lock_guard<std::mutex> lk(mu_);
  return shards_.size();
}

size_t GpuServer::state_bytes() {
  std::lock_guard<std::mutex> lk(mu_);
  size_t b = 0;
  for (auto& kv : shards_) b += kv.second.n * 12;
  return b;
}

size_t GpuServer::ElemsOf(const KVMeta& req) {
  const size_t bytes = static_cast<size_t>(req.val_len);
  if (req.cmd == kCmdInitF32) return bytes / 4;
  if (req.cmd == kCmdInitBf16) return bytes / 2;
  switch (req.codec) {
    case kCodecF32ToBf16:
    case kCodecF32ToFp8Block:
      return bytes / 4;
    default:
      return bytes / 2;  // bf16 sources (raw bf16, scaled bf16, bf16 -> fp8)
  }
}

int GpuServer::FormatOf(const KVMeta& req, int raw_format) {
  switch (req.codec) {
    case kCodecF32ToBf16:
    case kCodecBf16Scale:
      return PS_GRAD_BF16;
    case kCodecF32ToFp8Block:
    case kCodecBf16ToFp8Block:
      return PS_GRAD_FP8BLOCK;
    default:
      return raw_format;
  }
}

GpuServer::Shard* GpuServer::GetShard(Key key, size_t n) {
  auto it = shards_.find(key);
  if (it != shards_.end()) {
    if (n) CHECK_EQ(it->second.n, n) << "key " << key << " changed size";
    return &it->second;
  }
  CHECK_GT(n, (size_t)0) << "pull of unknown key " << key;
  Shard& s = shards_[key];
  s.n = n;
  GS_CUDA(cudaMalloc(&s.master, n * 4));
  GS_CUDA(cudaMalloc(&s.m, n * 4));
  GS_CUDA(cudaMalloc(&s.v, n * 4));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GS_CUDA(cudaMemsetAsync(s.master, 0, n * 4, st));
  GS_CUDA(cudaMemsetAsync(s.m, 0, n * 4, st));
  GS_CUDA(cudaMemsetAsync(s.v, 0, n * 4, st));
  s.slots.assign(cfg_.num_workers, nullptr);
  s.pushed.assign(cfg_.num_workers, 0);
  return &s;
}

void GpuServer::Handle(const KVMeta& req, const KVPairs<char>& data, KVServer<char>* /*server*/) {
  GS_CUDA(cudaSetDevice(dev_));
  std::lock_guard<std::mutex> lk(mu_);
  const Key key = data.keys.size() ? data.keys[0] : req.key;
  if (req.push) {
    Shard* s = GetShard(key, ElemsOf(req));
    if (req.cmd == kCmdInitBf16 || req.cmd == kCmdInitF32) {
      HandleInit(s, req, data);
    } else {
      HandleGrad(s, req, data);
      MaybeRunRound(key, s);
    }
  } else {
    Shard* s = GetShard(key, 0);
    HandlePull(s, req, data);
    MaybeRunRound(key, s);
  }
}

void GpuServer::HandleInit(Shard* s, const KVMeta& req, const KVPairs<char>& data) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (!s->initialized) {
    const bool f32 = req.cmd == kCmdInitF32;
    const void* src = data.vals.data();
    void* staged = nullptr;
    if (!data.vals.on_gpu()) {
      GS_CUDA(cudaMalloc(&staged, data.vals.size()));
      GS_CUDA(cudaMemcpyAsync(staged, src, data.vals.size(), cudaMemcpyHostToDevice, st));
      src = staged;
    }
    CHECK_EQ(ps_launch_decode(s->master, src, s->n, f32 ? PS_GRAD_F32 : PS_GRAD_BF16,
                              reinterpret_cast<ps_stream_t>(stream_)), 0);
    if (staged) {
      GS_CUDA(cudaStreamSynchronize(st));
      cudaFree(staged);
    }
    s->initialized = true;
  }
  server_->Response(req);
}

void GpuServer::HandleGrad(Shard* s, const KVMeta& req, const KVPairs<char>& data) {
  const int rank = Postoffice::IDtoRank(req.sender);
  CHECK_LT(rank, cfg_.num_workers);
  CHECK(!s->pushed[rank]) << "worker " << rank << " pushed key " << req.key
                          << " twice in one round";
  int fmt;
  const void* slot;
  if (req.mem.region == kSymmetricRegion) {
    // staged in the worker's symmetric gradient buffer: read later through the multicast address
    CHECK(mc_grad_base_ != nullptr) << "symmetric push but SetSymmetricGrads was never called";
    CHECK_EQ(FormatOf(req, PS_GRAD_BF16), (int)PS_GRAD_BF16) << "in-switch reduction adds bf16 values";
    CHECK_LE(req.mem.offset + s->n * 2, symm_grad_bytes_);
    CHECK_EQ(req.mem.offset % 16, (uint64_t)0);
    fmt = PS_GRAD_MC_BF16;
    slot = static_cast<const char*>(mc_grad_base_) + req.mem.offset;
  } else {
    CHECK(data.vals.on_gpu()) << "gradient pushes must arrive through the one-sided van";
    fmt = FormatOf(req, cfg_.raw_grad_format);
    slot = data.vals.data();
  }
  if (s->num_pushed == 0) s->grad_format = fmt;
  CHECK_EQ(s->grad_format, fmt) << "workers disagree on the gradient wire format";
  if (fmt == PS_GRAD_MC_BF16 && s->num_pushed > 0) {
    for (int w = 0; w < cfg_.num_workers; ++w) {
      if (s->pushed[w]) CHECK(s->slots[w] == slot) << "workers disagree on the symmetric offset";
    }
  }
  s->slots[rank] = slot;
  s->pushed[rank] = 1;
  ++s->num_pushed;
  // the payload already sits in its slot: the push is complete for the worker
  server_->Response(req);
}

void GpuServer::HandlePull(Shard* s, const KVMeta& req, const KVPairs<char>& /*data*/) {
  const int rank = Postoffice::IDtoRank(req.sender);
  CHECK_LT(rank, cfg_.num_workers);
  if (s->pushed[rank]) {
    s->waiting_pulls.push_back(req);  // wants the parameters *after* this round's update
  } else {
    ServePullFromLocal(req.key, s, req);
  }
}

void* GpuServer::WorkerDest(const KVMeta& pull) {
  if (!pull.mem.valid()) return nullptr;
  const int rank = Postoffice::IDtoRank(pull.sender);
  if (pull.mem.region == kSymmetricRegion) {
    if (peer_bases_.empty() || pull.mem.offset + pull.mem.bytes > symm_bytes_) return nullptr;
    return static_cast<char*>(peer_bases_[rank]) + pull.mem.offset;
  }
  const int instance_id = po_->GroupWorkerRankToInstanceID(rank, instance_idx_);
  return po_->van()->ResolvePeerMem(instance_id, pull.mem);
}

void GpuServer::MaybeRunRound(Key key, Shard* s) {
  const int W = cfg_.num_workers;
  if (s->num_pushed < W) return;
  if (cfg_.fuse_pull && static_cast<int>(s->waiting_pulls.size()) < W) return;

  ps_update_args a;
  memset(&a, 0, sizeof(a));
  a.n = s->n;
  a.num_grads = W;
  a.grad_format = s->grad_format;
  for (int w = 0; w < W; ++w) a.grads[w] = s->slots[w];
  if (s->grad_format == PS_GRAD_MC_BF16) {
    a.num_grads = 1;  // one stream: the switch returns the sum over all bound GPUs
    ++mc_reduce_;
  }
  a.master = s->master;
  a.m = s->m;
  a.v = s->v;
  a.num_outs = 0;  // no server-side bf16 copy: late pulls are cast from the fp32 master on demand
  std::vector<char> placed(s->waiting_pulls.size(), 0);
  bool all_symmetric = mc_base_ != nullptr && static_cast<int>(s->waiting_pulls.size()) == W;
  for (size_t i = 0; i < s->waiting_pulls.size(); ++i) {
    const KVMeta& pull = s->waiting_pulls[i];
    void* dst = WorkerDest(pull);
    if (dst && a.num_outs < PS_MAX_FANOUT && pull.mem.bytes >= s->n * 2 &&
        (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      a.outs[a.num_outs++] = dst;
      placed[i] = 1;
    }
    all_symmetric = all_symmetric && placed[i] && pull.mem.region == kSymmetricRegion &&
                    pull.mem.offset == s->waiting_pulls[0].mem.offset;
  }
  a.body_outs = a.num_outs;
  if (all_symmetric) {
    // every worker wants this shard at the same symmetric offset: one multicast store
    // stream replaces the W unicast streams (outs[1..] only serve the ragged tail)
    a.mc_out = static_cast<char*>(mc_base_) + s->waiting_pulls[0].mem.offset;
    a.body_outs = 0;
    ++mcast_;
  }
  ps_opt_params o = cfg_.opt;
  ++s->step;
  if (o.optimizer == PS_OPT_ADAMW) {
    o.bias_corr1 = 1.f - std::pow(o.beta1, static_cast<float>(s->step));
    o.bias_corr2 = 1.f - std::pow(o.beta2, static_cast<float>(s->step));
  }
  CHECK_EQ(ps_launch_update(&a, &o, cfg_.max_ctas, reinterpret_cast<ps_stream_t>(stream_)), 0);
  ++updates_;
  if (a.num_outs > 0) ++fused_;

  SendOpts placed_opts;
  placed_opts.codec = kCodecPlaced;
  for (size_t i = 0; i < s->waiting_pulls.size(); ++i) {
    const KVMeta& pull = s->waiting_pulls[i];
    if (placed[i]) {
      KVPairs<char> res;
      res.keys = OneKey(key);
      // the values are already in the worker's buffer; `vals` only conveys their size
      res.vals = SArray<char>(reinterpret_cast<char*>(s->master), s->n * 2, GPU, dev_, GPU, dev_);
      res.lens = OneLen(s->n * 2);
      server_->Response(pull, res, placed_opts);
    } else {
      ServePullFromLocal(key, s, pull);
    }
  }
  s->waiting_pulls.clear();
  std::fill(s->pushed.begin(), s->pushed.end(), 0);
  s->num_pushed = 0;
}

void GpuServer::ServePullFromLocal(Key key, Shard* s, const KVMeta& req) {
  KVPairs<char> res;
  res.keys = OneKey(key);
  res.lens = OneLen(s->n * 2);
  ps_stream_t st = reinterpret_cast<ps_stream_t>(stream_);
  if (void* dst = WorkerDest(req)) {
    // cast fp32 master -> bf16 straight into the worker's buffer (peer HBM); the kernel runs
    // on the update stream, so it is ordered after any update of this shard
    CHECK_GE(req.mem.bytes, s->n * 2) << "pull destination smaller than the shard";
    CHECK_EQ(ps_launch_copy(dst, s->master, s->n * 4, PS_CODEC_F32_TO_BF16, 1.f, cfg_.max_ctas, st),
             0);
    res.vals = SArray<char>(reinterpret_cast<char*>(s->master), s->n * 2, GPU, dev_, GPU, dev_);
    SendOpts placed;
    placed.codec = kCodecPlaced;
    server_->Response(req, res, placed);
    return;
  }
  CHECK(!req.mem.valid()) << "pull destination region unknown to this server";
  // two-sided requester: cast into a scratch buffer ...
  void* scratch = nullptr;
  GS_CUDA(cudaMalloc(&scratch, s->n * 2 + 16));
  CHECK_EQ(ps_launch_copy(scratch, s->master, s->n * 4, PS_CODEC_F32_TO_BF16, 1.f, cfg_.max_ctas, st),
           0);
  if (po_->van()->GetType() == "nccl") {
    // ... which the nccl van sends from device memory (it orders the send behind this stream)
    const int dev = dev_;
    SArray<char> dvals;
    dvals.reset(static_cast<char*>(scratch), s->n * 2,
                [dev](char* p) {
                  cudaSetDevice(dev);
                  cudaFree(p);
                },
                GPU, dev_, GPU, dev_);
    res.vals = dvals;
    server_->Response(req, res);
    return;
  }
  // ... or, for a host-only transport (TCP), stage it through host memory
  SArray<char> host(s->n * 2);
  cudaStream_t cst = static_cast<cudaStream_t>(stream_);
  GS_CUDA(cudaMemcpyAsync(host.data(), scratch, s->n * 2, cudaMemcpyDeviceToHost, cst));
  GS_CUDA(cudaStreamSynchronize(cst));
  cudaFree(scratch);
  res.vals = host;
  server_->Response(req, res);
}

bool GpuServer::ReadMaster(Key key, std::vector<float>* out) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = shards_.find(key);
  if (it == shards_.end()) return false;
  GS_CUDA(cudaSetDevice(dev_));
  out->resize(it->second.n);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GS_CUDA(cudaMemcpyAsync(out->data(), it->second.master, it->second.n * 4, cudaMemcpyDeviceToHost,
                          st));
  GS_CUDA(cudaStreamSynchronize(st));
  return true;
}

namespace {
struct CkptHeader {
  char magic[8];
  uint64_t num_keys;
};
struct CkptEntry {
  uint64_t key;
  uint64_t n;
  int32_t step;
  int32_t initialized;
};
}  // namespace

bool GpuServer::SaveCheckpoint(const std::string& path) {
  std::lock_guard<std::mutex> lk(mu_);
  GS_CUDA(cudaSetDevice(dev_));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GS_CUDA(cudaStreamSynchronize(st));
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  CkptHeader h;
  memcpy(h.magic, "PSB2CKPT", 8);
  h.num_keys = shards_.size();
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  std::vector<float> buf;
  for (auto& kv : shards_) {
    const Shard& s = kv.second;
    CkptEntry e{kv.first, s.n, s.step, s.initialized ? 1 : 0};
    ok = ok && fwrite(&e, sizeof(e), 1, f) == 1;
    buf.resize(s.n);
    for (float* src : {s.master, s.m, s.v}) {
      GS_CUDA(cudaMemcpy(buf.data(), src, s.n * 4, cudaMemcpyDeviceToHost));
      ok = ok && fwrite(buf.data(), 4, s.n, f) == s.n;
    }
  }
  fclose(f);
  return ok;
}

bool GpuServer::LoadCheckpoint(const std::string& path) {
  std::lock_guard<std::mutex> lk(mu_);
  GS_CUDA(cudaSetDevice(dev_));
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  CkptHeader h;
  bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "PSB2CKPT", 8) == 0;
  std::vector<float> buf;
  for (uint64_t i = 0; ok && i < h.num_keys; ++i) {
    CkptEntry e;
    ok = fread(&e, sizeof(e), 1, f) == 1;
    if (!ok) break;
    Shard* s = GetShard(e.key, e.n);
    s->step = e.step;
    s->initialized = e.initialized != 0;
    buf.resize(e.n);
    for (float* dst : {s->master, s->m, s->v}) {
      ok = ok && fread(buf.data(), 4, e.n, f) == e.n;
      GS_CUDA(cudaMemcpy(dst, buf.data(), e.n * 4, cudaMemcpyHostToDevice));
    }
  }
  GS_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream_)));
  fclose(f);
  return ok;
}

}  // namespace ps
