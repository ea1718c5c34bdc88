/**
 * \file fd_exchange.h
 * \brief FdExchange: a per-process rendezvous point for file descriptors and small values.
 *
 * CUDA's virtual-memory-management allocations (cuMemCreate) and NVSwitch multicast objects
 * (cuMulticastCreate) are shared between processes as POSIX file descriptors, and a file
 * descriptor can only cross a process boundary over a unix socket (SCM_RIGHTS). Every process
 * of a job therefore listens on an abstract unix socket; peers connect, name a key and receive
 * what was published under it. A fetch BLOCKS until the key exists, so the same mechanism is
 * the barrier of a collective allocation ("fetch token X from everyone" = everyone reached X).
 *
 * This is the bootstrap half of what the reference does with rkey / address exchange over its
 * rdma_cm channel (src/rdma_van.h:609-709) and with named shm segments for IPC
 * (src/rdma_transport.h:469-523).
 */
#ifndef PS_VAN_FD_EXCHANGE_H_
#define PS_VAN_FD_EXCHANGE_H_
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ps {

class FdExchange {
 public:
  /*! \brief the endpoint name of process `pid` in the job whose scheduler listens on `job_port` */
  static std::string EndpointName(int job_port, int pid) {
    return "pslite_b200_fdx_" + std::to_string(job_port) + "_" + std::to_string(pid);
  }

  /*! \brief process-wide instance for one job (created and started on first use) */
  static FdExchange* Get(int job_port) {
    static std::mutex mu;
    static std::map<int, std::unique_ptr<FdExchange>> all;
    std::lock_guard<std::mutex> lk(mu);
    auto& slot = all[job_port];
    if (!slot) {
      slot.reset(new FdExchange());
      if (!slot->Start(EndpointName(job_port, static_cast<int>(getpid())))) slot.reset();
    }
    return slot.get();
  }

  ~FdExchange() { Stop(); }

  /*! \brief make (fd, value) available under `key`; fd may be -1. The fd stays owned by the caller. */
  void Publish(const std::string& key, int fd, uint64_t value) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      Entry& e = entries_[key];
      e.fd = fd;
      e.value = value;
    }
    cv_.notify_all();
  }

  /*! \brief forget everything published under keys that start with `prefix` */
  void Retract(const std::string& prefix) {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto it = entries_.begin(); it != entries_.end();) {
      it = it->first.compare(0, prefix.size(), prefix) == 0 ? entries_.erase(it) : std::next(it);
    }
  }

  /*!
   * \brief get what the process listening on `endpoint` published under `key`, waiting up to
   *        `timeout_s` for the endpoint to exist and for the key to be published. On success
   *        *fd is a new descriptor owned by the caller (or -1 if none was published).
   */
  static bool Fetch(const std::string& endpoint, const std::string& key, int* fd, uint64_t* value,
                    int timeout_s = 120) {
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(timeout_s);
    int s = -1;
    for (;;) {
      s = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
      if (s < 0) return false;
      struct sockaddr_un sa;
      const socklen_t len = Addr(endpoint, &sa);
      if (connect(s, reinterpret_cast<struct sockaddr*>(&sa), len) == 0) break;
      close(s);
      if (std::chrono::steady_clock::now() > deadline) return false;
      std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    bool ok = false;
    const uint32_t klen = static_cast<uint32_t>(key.size());
    const int32_t wait_s = timeout_s;
    if (SendAll(s, &klen, sizeof(klen)) && SendAll(s, &wait_s, sizeof(wait_s)) && SendAll(s, key.data(), key.size())) {
      // reply: one message with {status, value} and, optionally, the descriptor as ancillary data
      struct {
        uint32_t status;
        uint32_t has_fd;
        uint64_t value;
      } rep;
      char ctrl[CMSG_SPACE(sizeof(int))];
      struct iovec iov = {&rep, sizeof(rep)};
      struct msghdr mh;
      memset(&mh, 0, sizeof(mh));
      mh.msg_iov = &iov;
      mh.msg_iovlen = 1;
      mh.msg_control = ctrl;
      mh.msg_controllen = sizeof(ctrl);
      struct pollfd pfd = {s, POLLIN, 0};
      const int ms = static_cast<int>(std::chrono::duration_cast<std::chrono::milliseconds>(
                                          deadline - std::chrono::steady_clock::now()).count());
      if (poll(&pfd, 1, ms > 0 ? ms + 1000 : 1000) > 0) {
        const ssize_t n = recvmsg(s, &mh, MSG_CMSG_CLOEXEC);
        if (n == static_cast<ssize_t>(sizeof(rep)) && rep.status == 1) {
          int got = -1;
          for (struct cmsghdr* c = CMSG_FIRSTHDR(&mh); c; c = CMSG_NXTHDR(&mh, c)) {
            if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) memcpy(&got, CMSG_DATA(c), sizeof(int));
          }
          if (!rep.has_fd || got >= 0) {
            if (fd) *fd = got; else if (got >= 0) close(got);
            if (value) *value = rep.value;
            ok = true;
          }
        }
      }
    }
    close(s);
    return ok;
  }

 private:
  struct Entry {
    int fd = -1;
    uint64_t value = 0;
  };

  static socklen_t Addr(const std::string& name, struct sockaddr_un* sa) {
    memset(sa, 0, sizeof(*sa));
    sa->sun_family = AF_UNIX;
    sa->sun_path[0] = '\0';  // abstract namespace: nothing to unlink, gone with the process
    const size_t n = std::min(name.size(), sizeof(sa->sun_path) - 2);
    memcpy(sa->sun_path + 1, name.data(), n);
    return static_cast<socklen_t>(offsetof(struct sockaddr_un, sun_path) + 1 + n);
  }
  static bool SendAll(int s, const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) {
      const ssize_t w = send(s, c, n, MSG_NOSIGNAL);
      if (w <= 0) {
        if (w < 0 && errno == EINTR) continue;
        return false;
      }
      c += w;
      n -= static_cast<size_t>(w);
    }
    return true;
  }
  static bool RecvAll(int s, void* p, size_t n) {
    char* c = static_cast<char*>(p);
    while (n) {
      const ssize_t r = recv(s, c, n, 0);
      if (r <= 0) {
        if (r < 0 && errno == EINTR) continue;
        return false;
      }
      c += r;
      n -= static_cast<size_t>(r);
    }
    return true;
  }

  bool Start(const std::string& endpoint) {
    const int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return false;
    struct sockaddr_un sa;
    const socklen_t len = Addr(endpoint, &sa);
    if (bind(fd, reinterpret_cast<struct sockaddr*>(&sa), len) != 0 || listen(fd, 256) != 0) {
      close(fd);
      return false;
    }
    listen_fd_.store(fd);
    acceptor_.reset(new std::thread([this, fd] { AcceptLoop(fd); }));
    return true;
  }

  void Stop() {
    stop_.store(true);
    cv_.notify_all();
    const int fd = listen_fd_.exchange(-1);
    if (fd >= 0) {
      shutdown(fd, SHUT_RDWR);
      close(fd);
    }
    if (acceptor_) acceptor_->join();
    acceptor_.reset();
    while (active_.load() > 0) std::this_thread::yield();
  }

  void AcceptLoop(int listen_fd) {
    for (;;) {
      const int c = accept4(listen_fd, nullptr, nullptr, SOCK_CLOEXEC);
      if (c < 0) {
        if (stop_.load()) return;
        if (errno == EINTR) continue;
        return;
      }
      // a request may have to wait for its key: serve each on its own (short-lived) thread
      active_.fetch_add(1);
      std::thread([this, c] {
        Serve(c);
        close(c);
        active_.fetch_sub(1);
      }).detach();
    }
  }

  void Serve(int c) {
    // descriptors of device memory are only handed to processes of the same user (abstract sockets
    // have no file permissions: anybody in the network namespace can connect)
    struct ucred cred;
    socklen_t cl = sizeof(cred);
    if (getsockopt(c, SOL_SOCKET, SO_PEERCRED, &cred, &cl) != 0 || cred.uid != geteuid()) return;
    uint32_t klen = 0;
    int32_t wait_s = 0;
    if (!RecvAll(c, &klen, sizeof(klen)) || !RecvAll(c, &wait_s, sizeof(wait_s)) || klen > 4096) return;
    std::string key(klen, '\0');
    if (klen && !RecvAll(c, &key[0], klen)) return;
    Entry e;
    bool found = false;
    {
      std::unique_lock<std::mutex> lk(mu_);
      found = cv_.wait_for(lk, std::chrono::seconds(wait_s > 0 ? wait_s : 1), [&] {
        return stop_.load() || entries_.count(key) > 0;
      }) && entries_.count(key) > 0;
      if (found) e = entries_[key];
    }
    struct {
      uint32_t status;
      uint32_t has_fd;
      uint64_t value;
    } rep = {found ? 1u : 0u, (found && e.fd >= 0) ? 1u : 0u, e.value};
    char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    struct iovec iov = {&rep, sizeof(rep)};
    struct msghdr mh;
    memset(&mh, 0, sizeof(mh));
    mh.msg_iov = &iov;
    mh.msg_iovlen = 1;
    if (rep.has_fd) {
      mh.msg_control = ctrl;
      mh.msg_controllen = sizeof(ctrl);
      struct cmsghdr* cm = CMSG_FIRSTHDR(&mh);
      cm->cmsg_level = SOL_SOCKET;
      cm->cmsg_type = SCM_RIGHTS;
      cm->cmsg_len = CMSG_LEN(sizeof(int));
      memcpy(CMSG_DATA(cm), &e.fd, sizeof(int));
    }
    ssize_t w;
    do {
      w = sendmsg(c, &mh, MSG_NOSIGNAL);
    } while (w < 0 && errno == EINTR);
  }

  std::atomic<int> listen_fd_{-1};
  std::unique_ptr<std::thread> acceptor_;
  std::atomic<bool> stop_{false};
  std::atomic<int> active_{0};
  std::mutex mu_;
  std::condition_variable cv_;
  std::map<std::string, Entry> entries_;
};

}  // namespace ps
#endif  // PS_VAN_FD_EXCHANGE_H_
