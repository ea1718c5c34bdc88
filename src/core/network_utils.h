/**
 * \file network_utils.h
 * \brief Interface / address / free-port discovery for the TCP control plane.
 * Parity: reference src/network_utils.h:28-264 (GetIP, GetAvailableInterfaceAndIP,
 * GetAvailablePort). Linux only — the B200 stack has no Windows target.
 */
#ifndef PS_CORE_NETWORK_UTILS_H_
#define PS_CORE_NETWORK_UTILS_H_
#include <arpa/inet.h>
#include <ifaddrs.h>
#include <net/if.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>
#include <array>
#include <cstring>
#include <string>
#include <vector>

namespace ps {

/*! \brief IPv4 address of `interface`, "" if it has none */
inline void GetIP(const std::string& interface, std::string* ip) {
  ip->clear();
  struct ifaddrs* list = nullptr;
  if (getifaddrs(&list) != 0) return;
  for (struct ifaddrs* it = list; it; it = it->ifa_next) {
    if (!it->ifa_addr || it->ifa_addr->sa_family != AF_INET) continue;
    if (interface != it->ifa_name) continue;
    char buf[INET_ADDRSTRLEN];
    auto* sin = reinterpret_cast<struct sockaddr_in*>(it->ifa_addr);
    if (inet_ntop(AF_INET, &sin->sin_addr, buf, sizeof(buf))) *ip = buf;
    break;
  }
  freeifaddrs(list);
}

/*! \brief first interface that is up, not loopback, and has an IPv4 address */
inline void GetAvailableInterfaceAndIP(std::string* interface, std::string* ip) {
  interface->clear();
  ip->clear();
  struct ifaddrs* list = nullptr;
  if (getifaddrs(&list) != 0) return;
  for (struct ifaddrs* it = list; it; it = it->ifa_next) {
    if (!it->ifa_addr || it->ifa_addr->sa_family != AF_INET) continue;
    if ((it->ifa_flags & IFF_LOOPBACK) || !(it->ifa_flags & IFF_UP)) continue;
    char buf[INET_ADDRSTRLEN];
    auto* sin = reinterpret_cast<struct sockaddr_in*>(it->ifa_addr);
    if (!inet_ntop(AF_INET, &sin->sin_addr, buf, sizeof(buf))) continue;
    *interface = it->ifa_name;
    *ip = buf;
    break;
  }
  freeifaddrs(list);
}

/*! \brief ask the kernel for one currently-free TCP port (0 on failure) */
inline int GetAvailablePort() {
  int fd = socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) return 0;
  struct sockaddr_in addr;
  memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_ANY);
  addr.sin_port = 0;
  int port = 0;
  if (bind(fd, reinterpret_cast<struct sockaddr*>(&addr), sizeof(addr)) == 0) {
    socklen_t len = sizeof(addr);
    if (getsockname(fd, reinterpret_cast<struct sockaddr*>(&addr), &len) == 0) {
      port = ntohs(addr.sin_port);
    }
  }
  close(fd);
  return port;
}

/*!
 * \brief reserve `num_ports` distinct free ports (sockets are held open until all
 *        are found so the kernel cannot hand out the same port twice)
 * \return how many were found
 */
inline int GetAvailablePort(int num_ports, std::array<int, 32>* ports) {
  std::vector<int> fds;
  int found = 0;
  for (int i = 0; i < num_ports && i < 32; ++i) {
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) break;
    struct sockaddr_in addr;
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = htonl(INADDR_ANY);
    if (bind(fd, reinterpret_cast<struct sockaddr*>(&addr), sizeof(addr)) != 0) {
      close(fd);
      break;
    }
    socklen_t len = sizeof(addr);
    getsockname(fd, reinterpret_cast<struct sockaddr*>(&addr), &len);
    (*ports)[i] = ntohs(addr.sin_port);
    fds.push_back(fd);
    ++found;
  }
  for (int fd : fds) close(fd);
  return found;
}

}  // namespace ps
#endif  // PS_CORE_NETWORK_UTILS_H_
