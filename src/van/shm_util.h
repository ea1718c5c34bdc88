/**
 * \file shm_util.h
 * \brief housekeeping for the POSIX shared-memory objects this library creates.
 *
 * Every object is named <prefix>n<pid-namespace>_<pid>_...; a process that is killed cannot
 * unlink its own, so each start sweeps the objects whose creator no longer exists (the
 * reference leaks its /dev/shm segments the same way: BytePS_ShM_* in src/rdma_van.h are never
 * unlinked). A pid only means something inside its pid namespace: with a /dev/shm shared between
 * containers (--ipc=host) kill(pid, 0) == ESRCH says nothing about another container's
 * process, so the namespace id is part of the name and only the caller's own objects are swept.
 */
#ifndef PS_VAN_SHM_UTIL_H_
#define PS_VAN_SHM_UTIL_H_
#include <dirent.h>
#include <signal.h>
#include <sys/mman.h>

#include <sys/stat.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>

namespace ps {

/*! \brief `base` + "n<inode of this process's pid namespace>_" (0 where /proc is unavailable) */
inline std::string ShmScopedPrefix(const char* base) {
  static const unsigned long ns = [] {
    struct stat st;
    return stat("/proc/self/ns/pid", &st) == 0 ? static_cast<unsigned long>(st.st_ino) : 0ul;
  }();
  return std::string(base) + "n" + std::to_string(ns) + "_";
}

/*! \brief unlink /dev/shm/<scoped prefix><pid>_* whose <pid> is gone; returns how many were removed */
inline int SweepStaleShm(const char* base) {
  DIR* d = opendir("/dev/shm");
  if (!d) return 0;
  const std::string scoped = ShmScopedPrefix(base);
  const char* prefix = scoped.c_str();
  const size_t plen = scoped.size();
  int removed = 0;
  while (struct dirent* e = readdir(d)) {
    if (strncmp(e->d_name, prefix, plen) != 0) continue;
    char* end = nullptr;
    const long pid = strtol(e->d_name + plen, &end, 10);
    if (end == e->d_name + plen || pid <= 0) continue;
    if (kill(static_cast<pid_t>(pid), 0) == 0 || errno != ESRCH) continue;  // creator is alive
    const std::string name = std::string("/") + e->d_name;
    if (shm_unlink(name.c_str()) == 0) ++removed;
  }
  closedir(d);
  return removed;
}

}  // namespace ps
#endif  // PS_VAN_SHM_UTIL_H_
