// Unit tests: SArray, Range, Environment, queues, wire codec, allocators.
#include <sys/wait.h>
#include <condition_variable>
#include <mutex>
#include <thread>
#include "core/wire.h"
#include "ps/internal/parallel_sort.h"
#include "ps/internal/spin_mutex.h"
#include "ps/internal/parallel_kv_match.h"
#include "ps/internal/message.h"
#include "test_util.h"
#include "van/mem_domain.h"
#include "van/shm_pipe.h"

using namespace ps;

TEST(sarray_basic) {
  SArray<int> a(4, 7);
  CHECK_EQ(a.size(), (size_t)4);
  CHECK_EQ(a[3], 7);
  a.push_back(9);
  CHECK_EQ(a.back(), 9);
  SArray<int> seg = a.segment(1, 3);
  CHECK_EQ(seg.size(), (size_t)2);
  seg[0] = 42;
  CHECK_EQ(a[1], 42);  // zero-copy view
  a.resize(2);
  CHECK_EQ(a.size(), (size_t)2);
  SArray<int> il{1, 2, 3};
  CHECK_EQ(il[2], 3);
}

TEST(sarray_cast_and_placement) {
  SArray<float> f(8, 1.0f);
  f.src_device_type_ = GPU;
  f.src_device_id_ = 3;
  f.dst_device_type_ = GPU;
  f.dst_device_id_ = 5;
  SArray<char> bytes(f);
  CHECK_EQ(bytes.size(), (size_t)32);
  CHECK_EQ((int)bytes.src_device_type_, (int)GPU);
  CHECK_EQ(bytes.dst_device_id_, 5);
  SArray<float> back(bytes);
  CHECK_EQ(back.size(), (size_t)8);
  CHECK(back.data() == f.data());
  SArray<char> odd(7);
  EXPECT_THROW(SArray<float> bad(odd));
}

TEST(sarray_keepalive) {
  int released = 0;
  {
    SArray<char> outer;
    {
      SArray<char> a;
      char* p = new char[16];
      a.reset(p, 16, [&released](char* q) { delete[] q; ++released; });
      outer = a.segment(4, 8);
    }
    CHECK_EQ(released, 0);  // the segment keeps the buffer alive
  }
  CHECK_EQ(released, 1);
}

TEST(sarray_append_self) {
  SArray<int> a{1, 2, 3};
  a.append(a);
  CHECK_EQ(a.size(), (size_t)6);
  CHECK_EQ(a[5], 3);
}

TEST(sarray_compact) {
  for (size_t n : {size_t(1), size_t(8), size_t(16), size_t(17), size_t(1000)}) {
    SArray<uint32_t> a = SArray<uint32_t>::Compact(n);
    CHECK_EQ(a.size(), n);
    CHECK_EQ(reinterpret_cast<uintptr_t>(a.data()) % alignof(uint32_t), (uintptr_t)0);
    for (size_t i = 0; i < n; ++i) a[i] = static_cast<uint32_t>(i * 3 + 1);
    SArray<uint32_t> b = a;  // shares the block
    SArray<char> bytes(a);   // converting view keeps it alive too
    a.clear();
    for (size_t i = 0; i < n; ++i) CHECK_EQ(b[i], static_cast<uint32_t>(i * 3 + 1));
    CHECK_EQ(bytes.size(), n * 4);
  }
  CHECK(SArray<char>::Compact(0).empty());
}

TEST(find_range) {
  SArray<Key> a{1, 3, 5, 7, 9};
  Range r = FindRange(a, (Key)2, (Key)7);
  CHECK_EQ(r.begin(), (uint64_t)1);
  CHECK_EQ(r.end(), (uint64_t)3);
}

TEST(environment_override) {
  Environment::Get()->set("PS_TEST_KNOB", "17");
  CHECK_EQ(GetEnv("PS_TEST_KNOB", 0), 17);
  CHECK_EQ(GetEnv("PS_TEST_MISSING", 5), 5);
  setenv("PS_TEST_FROM_OS", "abc", 1);
  CHECK_EQ(GetEnvStr("PS_TEST_FROM_OS"), std::string("abc"));
}

TEST(spsc_queue) {
  SPSCQueue<int> q(8);
  std::thread prod([&] {
    for (int i = 0; i < 100000; ++i) {
      while (!q.try_push(i)) {
      }
    }
  });
  long long sum = 0;
  for (int got = 0; got < 100000;) {
    int v;
    if (q.try_pop(&v)) {
      CHECK_EQ(v, got);
      sum += v;
      ++got;
    }
  }
  prod.join();
  CHECK_EQ(sum, 100000LL * 99999 / 2);
}

TEST(threadsafe_queue_both_modes) {
  for (const char* mode : {"0", "1"}) {
    Environment::Get()->set("DMLC_LOCKLESS_QUEUE", mode);
    ThreadsafeQueue<int> q;
    std::vector<std::thread> prods;
    for (int t = 0; t < 4; ++t) {
      prods.emplace_back([&q, t] {
        for (int i = 0; i < 5000; ++i) q.Push(t * 5000 + i);
      });
    }
    long long sum = 0;
    for (int i = 0; i < 20000; ++i) {
      int v;
      q.WaitAndPop(&v);
      sum += v;
    }
    for (auto& t : prods) t.join();
    CHECK_EQ(sum, 20000LL * 19999 / 2);
  }
  Environment::Get()->set("DMLC_LOCKLESS_QUEUE", "0");
}

TEST(spin_mutex_and_condition) {
  // mutual exclusion under contention, and sleeping on it through condition_variable_any
  SpinMutex mu;
  long counter = 0;
  std::vector<std::thread> th;
  for (int t = 0; t < 4; ++t) {
    th.emplace_back([&] {
      for (int i = 0; i < 50000; ++i) {
        std::lock_guard<SpinMutex> lk(mu);
        ++counter;
      }
    });
  }
  for (auto& t : th) t.join();
  CHECK_EQ(counter, 200000L);
  CHECK(mu.try_lock());
  CHECK(!mu.try_lock());
  mu.unlock();
  std::condition_variable_any cv;
  bool ready = false;
  std::thread waker([&] {
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
    {
      std::lock_guard<SpinMutex> lk(mu);
      ready = true;
    }
    cv.notify_all();
  });
  {
    std::unique_lock<SpinMutex> lk(mu);
    cv.wait(lk, [&] { return ready; });
  }
  waker.join();
}

TEST(spin_budget_and_poll) {
  // the window follows the long gaps: short ones (inside a burst) are ignored, gaps shorter than
  // the cap are polled through, gaps longer than the cap end the polling
  SpinBudget b(20, 1000);
  CHECK_EQ(b.window_us(), 1000);  // starts optimistic
  for (int i = 0; i < 100; ++i) b.Observe(3);  // inside a burst: no information
  CHECK_EQ(b.window_us(), 1000);
  for (int i = 0; i < 20; ++i) b.Observe(300);  // typical gap 300 us -> poll for twice that
  CHECK_GE(b.window_us(), 550);
  CHECK_LE(b.window_us(), 700);
  for (int i = 0; i < 20; ++i) b.Observe(50000);  // idle job: give up polling
  CHECK_EQ(b.window_us(), 20);
  for (int i = 0; i < 20; ++i) b.Observe(100);  // traffic is back
  CHECK_LE(b.window_us(), 400);
  CHECK_GE(b.window_us(), 150);
  SpinBudget fixed(50, 0);  // cap 0: fixed floor window
  fixed.Observe(10000);
  CHECK_EQ(fixed.window_us(), 50);

  std::atomic<bool> flag{false};
  CHECK(!SpinPoll([&] { return flag.load(); }, 10, 200));  // times out
  std::thread setter([&] {
    std::this_thread::sleep_for(std::chrono::microseconds(300));
    flag = true;
  });
  CHECK(SpinPoll([&] { return flag.load(); }, 10, 2000000));  // sees it in the polite phase
  setter.join();
  CHECK(SpinPoll([&] { return true; }, 0, 0));
}

TEST(wire_roundtrip_data_meta) {
  Meta m;
  m.head = 3; m.app_id = 1; m.customer_id = 2; m.timestamp = 77; m.request = true; m.push = true;
  m.body = "hello";
  m.data_type = {UINT64, CHAR, INT32};
  m.src_dev_type = GPU; m.src_dev_id = 1; m.dst_dev_type = GPU; m.dst_dev_id = 6;
  m.data_size = (int64_t)5 << 32;  // > 2 GiB must survive
  m.key = 0xfeedfacecafebeefULL; m.addr = 0x7f0000001000ULL; m.val_len = (int64_t)3 << 31;
  m.option = -2; m.sid = 9;
  m.mem.region = 4; m.mem.offset = 4096; m.mem.bytes = 1234567; m.mem.flag_seq = 11;
  m.codec = 3; m.scale = 0.125f;
  std::vector<char> buf;
  wire::PackMeta(m, &buf);
  CHECK_EQ(buf.size(), wire::PackedMetaSize(m));
  Meta r;
  CHECK(wire::UnpackMeta(buf.data(), buf.size(), &r));
  CHECK_EQ(r.head, 3); CHECK_EQ(r.timestamp, 77); CHECK(r.request); CHECK(r.push); CHECK(!r.simple_app);
  CHECK_EQ(r.body, std::string("hello"));
  CHECK_EQ(r.data_type.size(), (size_t)3); CHECK_EQ((int)r.data_type[2], (int)INT32);
  CHECK_EQ(r.data_size, (int64_t)5 << 32); CHECK_EQ(r.key, 0xfeedfacecafebeefULL);
  CHECK_EQ(r.val_len, (int64_t)3 << 31); CHECK_EQ(r.option, -2);
  CHECK_EQ(r.mem.region, 4); CHECK_EQ(r.mem.bytes, (uint64_t)1234567);
  CHECK_EQ(r.codec, 3); CHECK_EQ(r.scale, 0.125f);
  CHECK_EQ((int)r.dst_dev_type, (int)GPU); CHECK_EQ(r.dst_dev_id, 6);
  CHECK(!r.pull);
  CHECK(!r.pull_mem.valid());
  // truncated buffers are rejected, never mis-parsed
  for (size_t cut : {(size_t)0, (size_t)3, buf.size() / 2, buf.size() - 1}) {
    Meta bad;
    CHECK(!wire::UnpackMeta(buf.data(), cut, &bad));
  }
  // fused push-pull: the optional block travels only when the flag is set
  const size_t plain = buf.size();
  m.pull = true;
  m.pull_addr = 0x7f00deadb000ULL;
  m.pull_len = (int64_t)7 << 30;
  m.pull_mem.region = 0x40000000; m.pull_mem.offset = 1 << 20; m.pull_mem.bytes = 99;
  wire::PackMeta(m, &buf);
  CHECK_GT(buf.size(), plain);
  Meta p;
  CHECK(wire::UnpackMeta(buf.data(), buf.size(), &p));
  CHECK(p.push); CHECK(p.pull);
  CHECK_EQ(p.pull_addr, 0x7f00deadb000ULL); CHECK_EQ(p.pull_len, (int64_t)7 << 30);
  CHECK_EQ(p.pull_mem.region, 0x40000000); CHECK_EQ(p.pull_mem.offset, (uint64_t)1 << 20);
  CHECK_EQ(p.pull_mem.bytes, (uint64_t)99);
  CHECK_EQ(p.mem.region, 4);  // the push slot reference is independent of it
}

TEST(wire_roundtrip_control_nodes) {
  Meta m;
  m.control.cmd = Control::ADD_NODE;
  m.control.barrier_group = 7;
  m.control.msg_sig = 99;
  for (int i = 0; i < 3; ++i) {
    Node n;
    n.role = i ? Node::WORKER : Node::SERVER;
    n.id = 8 + i; n.hostname = "10.0.0." + std::to_string(i); n.port = 9000 + i;
    n.num_ports = 2; n.ports[0] = 9000 + i; n.ports[1] = 9100 + i;
    n.dev_types[1] = GPU; n.dev_ids[1] = i;
    n.is_recovery = i == 2; n.aux_id = i; n.pid = 1000 + i; n.dev_id = i;
    memcpy(n.endpoint_name, "0123456789abcdef", 16);
    n.endpoint_name_len = 16;
    m.control.node.push_back(n);
  }
  std::vector<char> buf;
  wire::PackMeta(m, &buf);
  Meta r;
  CHECK(wire::UnpackMeta(buf.data(), buf.size(), &r));
  CHECK_EQ((int)r.control.cmd, (int)Control::ADD_NODE);
  CHECK_EQ(r.control.node.size(), (size_t)3);
  CHECK_EQ(r.control.node[2].hostname, std::string("10.0.0.2"));
  CHECK(r.control.node[2].is_recovery);
  CHECK_EQ(r.control.node[1].ports[1], 9101);
  CHECK_EQ(r.control.node[1].dev_types[1], (int)GPU);
  CHECK_EQ(r.control.node[1].pid, 1001);
  CHECK_EQ(r.control.node[0].endpoint_name_len, (size_t)16);
  CHECK_EQ(memcmp(r.control.node[0].endpoint_name, "0123456789abcdef", 16), 0);
  // a node costs tens of bytes, not the reference's 552-byte RawNode
  CHECK_LT(buf.size(), (size_t)400);
}

TEST(arena_allocator) {
  ArenaAllocator a(1 << 20, 256);
  uint64_t x = a.Alloc(1000), y = a.Alloc(5000), z = a.Alloc(100);
  CHECK_EQ(x % 256, (uint64_t)0);
  CHECK(x != y && y != z);
  CHECK(a.Free(y));
  CHECK(!a.Free(y));
  uint64_t y2 = a.Alloc(4000);  // fits in the hole
  CHECK_EQ(y2, y);
  CHECK(a.Free(x)); CHECK(a.Free(y2)); CHECK(a.Free(z));
  CHECK_EQ(a.BytesInUse(), (uint64_t)0);
  CHECK_EQ(a.Alloc(1 << 20), (uint64_t)0);  // fully coalesced again
  CHECK_EQ(a.Alloc(1), UINT64_MAX);
}

TEST(shm_pipe_stream_and_doorbell) {
  const std::string name = "/pslite_test_pipe_" + std::to_string(getpid());
  auto tx = ShmPipe::Create(name, 8192);
  CHECK(tx != nullptr);
  auto rx = ShmPipe::Attach(name);
  CHECK(rx != nullptr);
  rx->Unlink();
  CHECK(ShmPipe::Attach(name) == nullptr);  // the name is gone, the mappings stay
  // a fresh ring starts "reader asleep": the very first frame must ring
  CHECK_EQ(rx->Readable(), (size_t)0);
  uint32_t hello = 0xabcd1234;
  CHECK(tx->Write(&hello, sizeof(hello)));
  CHECK(tx->ReaderNeedsDoorbell());
  CHECK(!tx->ReaderNeedsDoorbell());  // claimed exactly once
  uint32_t got = 0;
  CHECK(rx->Read(&got, sizeof(got)));
  CHECK_EQ(got, hello);
  // frames far larger than the 8 KB ring stream through it
  const size_t kFrame = 100000;
  const int kFrames = 20;
  std::thread producer([&] {
    std::vector<uint8_t> buf(kFrame);
    for (int f = 0; f < kFrames; ++f) {
      for (size_t i = 0; i < kFrame; ++i) buf[i] = static_cast<uint8_t>(i * 7 + f);
      CHECK(tx->Write(buf.data(), buf.size()));
    }
  });
  std::vector<uint8_t> in(kFrame);
  for (int f = 0; f < kFrames; ++f) {
    CHECK(rx->Read(in.data(), in.size()));
    for (size_t i = 0; i < kFrame; i += 997) CHECK_EQ(in[i], static_cast<uint8_t>(i * 7 + f));
  }
  producer.join();
  // sleep protocol: empty ring -> may sleep, writer then sees the flag; data present -> refuse
  CHECK(rx->PrepareSleep());
  CHECK(tx->Write(&hello, sizeof(hello)));
  CHECK(tx->ReaderNeedsDoorbell());
  CHECK(!rx->PrepareSleep());
  CHECK(rx->Read(&got, sizeof(got)));
}

TEST(shm_pipe_gate_and_peek) {
  // a gated frame is visible (Peek) but must not be consumed before the completion word says so
  const std::string name = "/pslite_test_gate_" + std::to_string(getpid());
  auto tx = ShmPipe::Create(name, 4096);
  auto rx = ShmPipe::Attach(name);
  CHECK(tx != nullptr && rx != nullptr);
  rx->Unlink();
  CHECK_EQ(rx->gate_done(), (uint64_t)0);
  struct Hdr { uint32_t magic; uint32_t pad; uint64_t gate; } h = {0x1234u, 0, 1}, peeked = {0, 0, 0};
  CHECK(!rx->Peek(&peeked, sizeof(peeked)));  // nothing published yet
  // wrap the header around the end of the ring: Peek must stitch it like Read does
  std::vector<char> filler(4096 - 8);
  CHECK(tx->Write(filler.data(), filler.size()));
  CHECK(rx->Read(filler.data(), filler.size()));
  rx->Commit();
  CHECK(tx->Write(&h, sizeof(h)));
  CHECK(rx->Peek(&peeked, sizeof(peeked)));
  CHECK_EQ(peeked.magic, 0x1234u);
  CHECK_EQ(peeked.gate, (uint64_t)1);
  CHECK_EQ(rx->Readable(), sizeof(h));        // peeking consumed nothing
  CHECK(rx->gate_done() < peeked.gate);       // gate closed
  // the "copy engine": another thread stores the completion through the word's address
  std::thread engine([&] {
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
    static_cast<std::atomic<uint64_t>*>(tx->gate_word())->store(1, std::memory_order_release);
  });
  while (rx->gate_done() < peeked.gate) std::this_thread::yield();
  engine.join();
  Hdr got = {0, 0, 0};
  CHECK(rx->Read(&got, sizeof(got)));
  CHECK_EQ(got.gate, (uint64_t)1);
  tx->SignalGate(2);
  CHECK_EQ(rx->gate_done(), (uint64_t)2);
}

TEST(shm_pipe_no_lost_wakeups) {
  // the reader really sleeps (on a condition variable standing in for the socket doorbell) and
  // must be woken for every frame that arrives while it sleeps: a lost wake-up shows as a timeout
  const std::string name = "/pslite_test_wake_" + std::to_string(getpid());
  auto tx = ShmPipe::Create(name, 4096);
  auto rx = ShmPipe::Attach(name);
  CHECK(tx != nullptr && rx != nullptr);
  rx->Unlink();
  std::mutex mu;
  std::condition_variable bell;
  int rings = 0;
  const int kFrames = 20000;
  std::atomic<bool> lost{false};
  std::thread writer([&] {
    unsigned seed = 7;
    for (int f = 0; f < kFrames; ++f) {
      uint32_t v = static_cast<uint32_t>(f);
      CHECK(tx->Write(&v, sizeof(v)));
      if (tx->ReaderNeedsDoorbell()) {
        std::lock_guard<std::mutex> lk(mu);
        ++rings;
        bell.notify_one();
      }
      if ((rand_r(&seed) & 127) == 0) std::this_thread::sleep_for(std::chrono::microseconds(rand_r(&seed) % 300));
    }
  });
  int seen_rings = 0, sleeps = 0;
  for (int f = 0; f < kFrames && !lost.load(); ++f) {
    while (rx->Readable() < sizeof(uint32_t)) {
      if (!rx->PrepareSleep()) continue;  // data showed up while announcing the sleep
      std::unique_lock<std::mutex> lk(mu);
      ++sleeps;
      if (!bell.wait_for(lk, std::chrono::seconds(5), [&] { return rings > seen_rings || rx->Readable() > 0; })) {
        lost.store(true);
        break;
      }
      seen_rings = rings;
      lk.unlock();
      rx->CancelSleep();
    }
    if (lost.load()) break;
    uint32_t v = 0;
    CHECK(rx->Read(&v, sizeof(v)));
    CHECK_EQ(v, static_cast<uint32_t>(f));
  }
  writer.join();
  CHECK(!lost.load()) << "a frame arrived while the reader slept and nobody rang";
  CHECK_GT(sleeps, 0);
}

TEST(stale_shm_sweep) {
  // an object named after a process that no longer exists is removed; a live creator's is kept
  pid_t child = fork();
  if (child == 0) _exit(0);
  int status = 0;
  waitpid(child, &status, 0);
  const std::string dead = "/" + ShmScopedPrefix("pslite_sweeptest_") + std::to_string(child) + "_0";
  const std::string live = "/" + ShmScopedPrefix("pslite_sweeptest_") + std::to_string(getpid()) + "_0";
  // same pid, but created in another pid namespace (another container sharing /dev/shm): not ours
  const std::string foreign = "/pslite_sweeptest_n1_" + std::to_string(child) + "_0";
  for (const std::string& n : {dead, live, foreign}) {
    int fd = shm_open(n.c_str(), O_CREAT | O_RDWR, 0600);
    CHECK_GE(fd, 0);
    close(fd);
  }
  CHECK_GE(SweepStaleShm("pslite_sweeptest_"), 1);
  CHECK_LT(shm_open(dead.c_str(), O_RDWR, 0600), 0);
  for (const std::string& n : {live, foreign}) {
    int fd = shm_open(n.c_str(), O_RDWR, 0600);
    CHECK_GE(fd, 0) << n << " was swept";
    close(fd);
    shm_unlink(n.c_str());
  }
}

TEST(index_pool) {
  IndexPool<int> pool(4);
  int v[6];
  uint32_t idx[6];
  for (int i = 0; i < 6; ++i) idx[i] = pool.Store(&v[i]);  // grows past the initial 4
  for (int i = 0; i < 6; ++i) CHECK(pool.Get(idx[i]) == &v[i]);
  CHECK(pool.Release(idx[2]) == &v[2]);
  CHECK(pool.Get(idx[2]) == nullptr);
}

TEST(wire_bytes_table) {
  CHECK_EQ(WireBytes(kCodecRaw, 1000), (uint64_t)1000);
  CHECK_EQ(WireBytes(kCodecF32ToBf16, 4000), (uint64_t)2000);
  CHECK_EQ(WireBytes(kCodecF32ToFp8Block, 4 * 64), (uint64_t)(64 + 2));
  CHECK_EQ(WireBytes(kCodecBf16ToFp8Block, 2 * 33), (uint64_t)(64 + 2));  // padded to 2 blocks
}

TEST(parallel_sort_and_match) {
  SArray<int> a(100000);
  for (size_t i = 0; i < a.size(); ++i) a[i] = (int)((i * 7919) % 100003);
  ParallelSort(&a, 4, std::less<int>());
  for (size_t i = 1; i < a.size(); ++i) CHECK_LE(a[i - 1], a[i]);
  SArray<Key> src_k{1, 3, 5, 7};
  SArray<float> src_v{1.f, 3.f, 5.f, 7.f};
  SArray<Key> dst_k{3, 4, 7, 9};
  SArray<float> dst_v(4, 10.f);
  size_t matched = ParallelOrderedMatch(src_k, src_v, dst_k, &dst_v, 1, PLUS, 2);
  CHECK_EQ(matched, (size_t)2);
  CHECK_EQ(dst_v[0], 13.f); CHECK_EQ(dst_v[1], 10.f); CHECK_EQ(dst_v[2], 17.f);
}

TEST(inline_vec_inline_and_spill) {
  // the first N elements live inside the object, more spill into a vector; copies are deep
  InlineVec<SArray<char>, 4> v;
  CHECK(v.empty());
  std::vector<SArray<char>> keep;
  for (int i = 0; i < 7; ++i) {
    SArray<char> a(static_cast<size_t>(i + 1), static_cast<char>('a' + i));
    keep.push_back(a);
    v.push_back(a);
    CHECK_EQ(v.size(), static_cast<size_t>(i + 1));
    CHECK_EQ(v.back().size(), static_cast<size_t>(i + 1));
  }
  size_t n = 0;
  for (const SArray<char>& a : v) {
    CHECK_EQ(a.size(), n + 1);
    CHECK_EQ(a[0], static_cast<char>('a' + n));
    ++n;
  }
  CHECK_EQ(n, (size_t)7);
  InlineVec<SArray<char>, 4> copy = v;
  v.resize(2);
  CHECK_EQ(v.size(), (size_t)2);
  CHECK_EQ(copy.size(), (size_t)7);
  CHECK_EQ(copy[6].size(), (size_t)7);
  v.clear();
  CHECK(v.empty());
  v.push_back(keep[3]);
  CHECK_EQ(v[0].data(), keep[3].data());  // a view, not a copy of the bytes
  InlineVec<DataType, 4> t = {CHAR, INT32};
  InlineVec<DataType, 4> u = {CHAR, INT32};
  CHECK(t == u);
  u.push_back(FLOAT);
  CHECK(t != u);
  // a Message with more segments than the inline capacity still round-trips
  Message m;
  for (int i = 0; i < 6; ++i) m.AddData(SArray<char>(static_cast<size_t>(8 + i), 'x'));
  CHECK_EQ(m.data.size(), (size_t)6);
  CHECK_EQ(m.meta.data_type.size(), (size_t)6);
  CHECK_EQ(m.data[5].size(), (size_t)13);
}

TEST(fd_exchange_publish_fetch_and_barrier) {
  // descriptors and small values between processes: the child publishes a pipe's write end under a
  // key, the parent fetches it (blocking until it exists) and writes through it
  const int job = 40000 + static_cast<int>(getpid() % 20000);
  int sync_pipe[2];
  CHECK_EQ(pipe(sync_pipe), 0);
  pid_t child = fork();
  if (child == 0) {
    int data_pipe[2];
    if (pipe(data_pipe) != 0) _exit(2);
    FdExchange* fx = FdExchange::Get(job);
    if (!fx) _exit(3);
    usleep(50 * 1000);  // the parent is already waiting for the key by now
    fx->Publish("tag/mem", data_pipe[1], 4242);
    char got[6] = {0};
    if (read(data_pipe[0], got, 5) != 5 || strcmp(got, "hello") != 0) _exit(4);
    // the "everybody is done" token: fetch the parent's, which it publishes after writing
    uint64_t v = 0;
    const std::string parent = FdExchange::EndpointName(job, static_cast<int>(getppid()));
    if (!FdExchange::Fetch(parent, "tag/done", nullptr, &v, 20) || v != 1) _exit(5);
    _exit(0);
  }
  FdExchange* fx = FdExchange::Get(job);
  CHECK(fx != nullptr);
  int fd = -1;
  uint64_t value = 0;
  CHECK(FdExchange::Fetch(FdExchange::EndpointName(job, static_cast<int>(child)), "tag/mem", &fd, &value, 20));
  CHECK_EQ(value, (uint64_t)4242);
  CHECK_GE(fd, 0);
  CHECK_EQ(write(fd, "hello", 5), (ssize_t)5);
  close(fd);
  fx->Publish("tag/done", -1, 1);
  int status = 0;
  waitpid(child, &status, 0);
  CHECK(WIFEXITED(status));
  CHECK_EQ(WEXITSTATUS(status), 0);
  // a key nobody publishes times out instead of hanging
  CHECK(!FdExchange::Fetch(FdExchange::EndpointName(job, static_cast<int>(getpid())), "tag/never", nullptr, &value, 1));
  fx->Retract("tag/");
  close(sync_pipe[0]);
  close(sync_pipe[1]);
}

TEST(logging_check_throws) {
  EXPECT_THROW(CHECK_EQ(1, 2) << "boom");
  EXPECT_THROW(LOG(FATAL) << "fatal");
  int x = 3;
  CHECK_NOTNULL(&x);
}

int main() { return RunAllTests(); }
