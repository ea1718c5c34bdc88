/**
 * \file wire.cc
 * \brief Meta / Node serialisation (see wire.h for the format rationale).
 */
#include "core/wire.h"

namespace ps {
namespace wire {

namespace {
enum MetaFlags : uint16_t {
  kFlagRequest = 1u << 0,
  kFlagPush = 1u << 1,
  kFlagSimpleApp = 1u << 2,
  kFlagHasControl = 1u << 3,
  kFlagHasMem = 1u << 4,
  kFlagPull = 1u << 5,
};
}  // namespace

void PackNode(const Node& n, Writer* w) {
  w->Put<uint8_t>(static_cast<uint8_t>(n.role));
  w->Put<int32_t>(n.id);
  w->Put<int32_t>(n.customer_id);
  w->PutString(n.hostname);
  int np = n.num_ports < 0 ? 0 : (n.num_ports > kMaxNodePorts ? kMaxNodePorts : n.num_ports);
  w->Put<uint8_t>(static_cast<uint8_t>(np));
  for (int i = 0; i < np; ++i) {
    w->Put<int32_t>(n.ports[i]);
    w->Put<uint8_t>(static_cast<uint8_t>(n.dev_types[i]));
    w->Put<int32_t>(n.dev_ids[i]);
  }
  w->Put<int32_t>(n.port);
  w->Put<uint8_t>(n.is_recovery ? 1 : 0);
  uint8_t el = static_cast<uint8_t>(n.endpoint_name_len > 64 ? 64 : n.endpoint_name_len);
  w->Put<uint8_t>(el);
  w->PutBytes(n.endpoint_name, el);
  w->Put<int32_t>(n.aux_id);
  w->Put<int32_t>(n.pid);
  w->Put<int32_t>(n.dev_id);
}

bool UnpackNode(Reader* r, Node* n) {
  n->role = static_cast<Node::Role>(r->Get<uint8_t>());
  n->id = r->Get<int32_t>();
  n->customer_id = r->Get<int32_t>();
  n->hostname = r->GetString();
  int np = r->Get<uint8_t>();
  if (np > kMaxNodePorts) return false;
  n->num_ports = np;
  for (int i = 0; i < np; ++i) {
    n->ports[i] = r->Get<int32_t>();
    n->dev_types[i] = r->Get<uint8_t>();
    n->dev_ids[i] = r->Get<int32_t>();
  }
  n->port = r->Get<int32_t>();
  n->is_recovery = r->Get<uint8_t>() != 0;
  uint8_t el = r->Get<uint8_t>();
  if (el > 64) return false;
  n->endpoint_name_len = el;
  r->GetBytes(n->endpoint_name, el);
  n->aux_id = r->Get<int32_t>();
  n->pid = r->Get<int32_t>();
  n->dev_id = r->Get<int32_t>();
  return r->ok();
}

void PackMeta(const Meta& m, std::vector<char>* out) {
  out->clear();
  out->resize(256 + m.body.size());  // one growth for the common descriptor; the cursor starts at 0
  Writer w(out, 0);
  w.Put<uint32_t>(kMetaMagic);
  w.Put<uint16_t>(kMetaVersion);
  uint16_t flags = 0;
  if (m.request) flags |= kFlagRequest;
  if (m.push) flags |= kFlagPush;
  if (m.simple_app) flags |= kFlagSimpleApp;
  if (!m.control.empty()) flags |= kFlagHasControl;
  if (m.mem.valid()) flags |= kFlagHasMem;
  if (m.pull) flags |= kFlagPull;
  w.Put<uint16_t>(flags);
  w.Put<int32_t>(m.head);
  w.Put<int32_t>(m.app_id);
  w.Put<int32_t>(m.customer_id);
  w.Put<int32_t>(m.timestamp);
  w.Put<uint8_t>(static_cast<uint8_t>(m.src_dev_type));
  w.Put<int32_t>(m.src_dev_id);
  w.Put<uint8_t>(static_cast<uint8_t>(m.dst_dev_type));
  w.Put<int32_t>(m.dst_dev_id);
  w.Put<int64_t>(m.data_size);
  w.Put<uint64_t>(m.key);
  w.Put<uint64_t>(m.addr);
  w.Put<int64_t>(m.val_len);
  w.Put<int32_t>(m.option);
  w.Put<int32_t>(m.sid);
  if (flags & kFlagHasMem) {
    w.Put<int32_t>(m.mem.region);
    w.Put<uint64_t>(m.mem.offset);
    w.Put<uint64_t>(m.mem.bytes);
    w.Put<uint64_t>(m.mem.flag_seq);
  }
  if (flags & kFlagPull) {
    w.Put<uint64_t>(m.pull_addr);
    w.Put<int64_t>(m.pull_len);
    w.Put<int32_t>(m.pull_mem.region);
    w.Put<uint64_t>(m.pull_mem.offset);
    w.Put<uint64_t>(m.pull_mem.bytes);
  }
  w.Put<int32_t>(m.codec);
  w.Put<float>(m.scale);
  w.PutString(m.body);
  w.Put<uint8_t>(static_cast<uint8_t>(m.data_type.size()));
  for (DataType d : m.data_type) w.Put<uint8_t>(static_cast<uint8_t>(d));
  if (flags & kFlagHasControl) {
    w.Put<uint8_t>(static_cast<uint8_t>(m.control.cmd));
    w.Put<int32_t>(m.control.barrier_group);
    w.Put<uint64_t>(m.control.msg_sig);
    w.Put<uint32_t>(static_cast<uint32_t>(m.control.node.size()));
    for (const Node& n : m.control.node) PackNode(n, &w);
  }
}

size_t PackedMetaSize(const Meta& m) {
  std::vector<char> tmp;
  PackMeta(m, &tmp);
  return tmp.size();
}

bool UnpackMeta(const char* buf, size_t len, Meta* m) {
  Reader r(buf, len);
  if (r.Get<uint32_t>() != kMetaMagic) return false;
  if (r.Get<uint16_t>() != kMetaVersion) return false;
  uint16_t flags = r.Get<uint16_t>();
  m->request = flags & kFlagRequest;
  m->push = flags & kFlagPush;
  m->pull = flags & kFlagPull;
  m->simple_app = flags & kFlagSimpleApp;
  m->head = r.Get<int32_t>();
  m->app_id = r.Get<int32_t>();
  m->customer_id = r.Get<int32_t>();
  m->timestamp = r.Get<int32_t>();
  m->src_dev_type = static_cast<DeviceType>(r.Get<uint8_t>());
  m->src_dev_id = r.Get<int32_t>();
  m->dst_dev_type = static_cast<DeviceType>(r.Get<uint8_t>());
  m->dst_dev_id = r.Get<int32_t>();
  m->data_size = r.Get<int64_t>();
  m->key = r.Get<uint64_t>();
  m->addr = r.Get<uint64_t>();
  m->val_len = r.Get<int64_t>();
  m->option = r.Get<int32_t>();
  m->sid = r.Get<int32_t>();
  m->mem = MemRef();
  if (flags & kFlagHasMem) {
    m->mem.region = r.Get<int32_t>();
    m->mem.offset = r.Get<uint64_t>();
    m->mem.bytes = r.Get<uint64_t>();
    m->mem.flag_seq = r.Get<uint64_t>();
  }
  m->pull_addr = 0;
  m->pull_len = 0;
  m->pull_mem = MemRef();
  if (flags & kFlagPull) {
    m->pull_addr = r.Get<uint64_t>();
    m->pull_len = r.Get<int64_t>();
    m->pull_mem.region = r.Get<int32_t>();
    m->pull_mem.offset = r.Get<uint64_t>();
    m->pull_mem.bytes = r.Get<uint64_t>();
  }
  m->codec = r.Get<int32_t>();
  m->scale = r.Get<float>();
  m->body = r.GetString();
  int ndt = r.Get<uint8_t>();
  m->data_type.resize(ndt);
  for (int i = 0; i < ndt; ++i) m->data_type[i] = static_cast<DataType>(r.Get<uint8_t>());
  m->control = Control();
  if (flags & kFlagHasControl) {
    m->control.cmd = static_cast<Control::Command>(r.Get<uint8_t>());
    m->control.barrier_group = r.Get<int32_t>();
    m->control.msg_sig = r.Get<uint64_t>();
    uint32_t nn = r.Get<uint32_t>();
    if (!r.ok() || nn > (1u << 20)) return false;
    m->control.node.resize(nn);
    for (uint32_t i = 0; i < nn; ++i) {
      if (!UnpackNode(&r, &m->control.node[i])) return false;
    }
  }
  return r.ok();
}

}  // namespace wire
}  // namespace ps
