// Compile-time check that the public API of the reference (include/ps/*.h of
// bytedance/ps-lite) is available with the same names and call shapes. Nothing here talks to a
// cluster: every expression only has to type-check (plus a few value checks on constants).
#include <type_traits>
#include "ps/ps.h"
#include "test_util.h"

using namespace ps;

namespace {
template <typename T>
void Use(T&&) {}

// never executed: bodies exist so that every call expression is instantiated by the compiler
void KVWorkerSurface(KVWorker<float>* w) {
  std::vector<Key> keys;
  std::vector<float> vals;
  std::vector<int> lens;
  KVWorker<float>::Callback cb = [] {};
  int ts = w->Push(keys, vals);
  ts = w->Push(keys, vals, lens, 7, cb);
  ts = w->Pull(keys, &vals);
  ts = w->Pull(keys, &vals, &lens, 7, cb);
  w->Wait(ts);
  SArray<Key> zk;
  SArray<float> zv;
  SArray<int> zl;
  ts = w->ZPush(zk, zv);
  ts = w->ZPush(zk, zv, zl, 7, cb);
  ts = w->ZPull(zk, &zv);
  ts = w->ZPull(zk, &zv, &zl, 7, cb);
  ts = w->ZPushPull(zk, zv, &zv);                 // extension: fused push + pull
  ts = w->PushPull(keys, vals, &vals);
  KVWorker<float>::Slicer slicer = [](const KVPairs<float>&, const std::vector<Range>&,
                                      KVWorker<float>::SlicedKVs*) {};
  w->set_slicer(slicer);
  // SimpleApp surface inherited by the KV apps
  ts = w->Request(1, "body", kServerGroup);
  Use(w->get_customer());
}

void KVServerSurface(KVServer<float>* s) {
  KVServer<float>::ReqHandle h = [](const KVMeta& m, const KVPairs<float>& d, KVServer<float>* self) {
    Use(m.cmd); Use(m.push); Use(m.sender); Use(m.timestamp); Use(m.customer_id); Use(m.key);
    Use(m.addr); Use(m.val_len); Use(m.option);
    self->Response(m);
    self->Response(m, d);
  };
  s->set_request_handle(h);
  s->set_request_handle(KVServerDefaultHandle<float>());
  SArray<Key> k;
  SArray<float> v;
  SArray<int> l;
  s->RegisterRecvBufferWithRank(0, k, v);
  s->RegisterRecvBufferWithRank(0, k, v, l, 3);
  Use(s->instance_idx_);
}

void SimpleAppSurface(SimpleApp* a) {
  SimpleApp::Handle h = [](const SimpleData& d, SimpleApp* app) {
    Use(d.head); Use(d.body); Use(d.sender); Use(d.timestamp); Use(d.customer_id);
    app->Response(d);
    app->Response(d, "reply");
  };
  a->set_request_handle(h);
  a->set_response_handle(h);
  a->Wait(a->Request(0, "", kScheduler));
}

void PostofficeSurface(Postoffice* po, Customer* c) {
  Use(Postoffice::Get()); Use(Postoffice::GetWorker(0)); Use(Postoffice::GetServer(0));
  Use(Postoffice::GetScheduler());
  Use(po->van()); Use(po->GetNodeIDs(kWorkerGroup)); Use(po->GetServerKeyRanges());
  Use(po->GetCustomer(0, 0)); Use(po->GetCustomer(0, 0, 5));
  po->AddCustomer(c); po->RemoveCustomer(c);
  Use(po->GroupWorkerRankToInstanceID(0, 0)); Use(po->GroupServerRankToInstanceID(0, 0));
  Use(po->InstanceIDtoGroupRank(9));
  Use(Postoffice::WorkerRankToID(0)); Use(Postoffice::ServerRankToID(0)); Use(Postoffice::IDtoRank(9));
  Use(po->group_size()); Use(po->num_workers()); Use(po->num_servers());
  Use(po->num_worker_instances()); Use(po->num_server_instances());
  Use(po->my_rank()); Use(po->preferred_rank());
  Use(po->is_worker()); Use(po->is_server()); Use(po->is_scheduler()); Use(po->role_str());
  Use(po->verbose()); Use(po->is_recovery());
  po->Barrier(0, kWorkerGroup);
  po->RegisterExitCallback([] {});
  Use(po->GetDeadNodes(60));
  Van* van = po->van();
  Message m;
  Use(van->Send(m)); Use(van->my_node()); Use(van->GetTimestamp()); Use(van->IsReady());
  Use(van->GetType());
  van->RegisterRecvBuffer(m);
  van->PinMemory(nullptr, 0, false);
  Use(c->app_id()); Use(c->customer_id());
  int ts = c->NewRequest(kServerGroup);
  c->WaitRequest(ts); Use(c->NumResponse(ts)); c->AddResponse(ts); c->AddResponse(ts, 2);
  c->Accept(m);
}
}  // namespace

TEST(base_constants) {
  static_assert(std::is_same<Key, uint64_t>::value || sizeof(Key) == 4, "Key is uint64_t (or 32-bit with USE_KEY32)");
  CHECK_EQ(static_cast<int>(kScheduler), 1);
  CHECK_EQ(static_cast<int>(kServerGroup), 2);
  CHECK_EQ(static_cast<int>(kWorkerGroup), 4);
  CHECK_GT(kMaxKey, static_cast<Key>(1) << 30);
  CHECK_EQ(static_cast<int>(UNK), 0);
  CHECK_EQ(static_cast<int>(CPU), 1);
  CHECK_EQ(static_cast<int>(GPU), 2);
}

TEST(free_functions_exist) {
  // ps.h: only their addresses are taken (calling them needs a running cluster)
  Use(&NumWorkers); Use(&NumServers); Use(&IsServer); Use(&IsScheduler); Use(&MyRank);
  Use(&GetRole);
  void (*start)(int, Node::Role, int, bool, const char*) = &StartPS;
  void (*fin)(int, Node::Role, bool) = &Finalize;
  void (*exitcb)(const std::function<void()>&) = &RegisterExitCallback;
  CHECK(start != nullptr && fin != nullptr && exitcb != nullptr);
  CHECK(GetRole("worker") == Node::WORKER);
  CHECK(GetRole("server") == Node::SERVER);
  CHECK(GetRole("scheduler") == Node::SCHEDULER);
  CHECK(GetRole("joint") == Node::JOINT);
}

TEST(sarray_surface) {
  SArray<float> a;
  SArray<float> b(4, 1.f);
  SArray<double> c(b);  // converting view
  float raw[3] = {1, 2, 3};
  SArray<float> d(raw, 3);
  SArray<float> e(raw, 3, false);
  SArray<float> f(raw, 3, GPU, 1, CPU, 0, false);
  CHECK_EQ(static_cast<int>(f.src_device_type_), static_cast<int>(GPU));
  CHECK_EQ(f.src_device_id_, 1);
  CHECK_EQ(static_cast<int>(f.dst_device_type_), static_cast<int>(CPU));
  std::vector<float> v{1, 2};
  SArray<float> g(v);
  SArray<float> h(std::make_shared<std::vector<float>>(v));
  a.CopyFrom(raw, 3);
  a.CopyFrom(b);
  a.reset(raw, 3, [](float*) {});
  a.reset(raw, 3, [](float*) {}, GPU, 0, GPU, 0);
  b.resize(8); b.reserve(16); b.push_back(2.f); b.pop_back(); b.append(g);
  CHECK_EQ(b.size(), (size_t)10);
  CHECK(!b.empty());
  Use(b.capacity()); Use(b.begin()); Use(b.end()); Use(b.data()); Use(b.ptr()); Use(b.back());
  Use(b.front()); Use(b[0]); Use(b.segment(0, 2)); Use(b.DebugString());
  b.clear();
  CHECK(b.empty());
}

TEST(kv_and_app_classes_type_check) {
  // taking the addresses forces the (never called) surface functions to be compiled
  Use(&KVWorkerSurface); Use(&KVServerSurface); Use(&SimpleAppSurface); Use(&PostofficeSurface);
  KVPairs<float> kv;
  Use(kv.keys); Use(kv.vals); Use(kv.lens);
  static_assert(std::is_base_of<SimpleApp, KVWorker<float>>::value, "KVWorker is a SimpleApp");
  static_assert(std::is_base_of<SimpleApp, KVServer<float>>::value, "KVServer is a SimpleApp");
  static_assert(std::is_constructible<KVWorker<char>, int, int>::value, "KVWorker(app, customer)");
  static_assert(std::is_constructible<KVWorker<char>, int, int, int>::value, "KVWorker(app, customer, instance)");
  static_assert(std::is_constructible<KVServer<char>, int>::value, "KVServer(app)");
  static_assert(std::is_constructible<KVServer<char>, int, bool, int>::value, "KVServer(app, is_scheduler, instance)");
  static_assert(std::is_constructible<SimpleApp, int, int, Postoffice*>::value, "SimpleApp(app, customer, po)");
  Use(static_cast<Van* (*)(const std::string&, Postoffice*)>(&Van::Create));
}

int main() { return RunAllTests(); }
