/**
 * \file shm_util.h
 * \brief housekeeping for the POSIX shared-memory objects this library creates.
 *
 * Every object is named <prefix><pid>_...; a process that is killed cannot unlink its own,
 * so each start sweeps the objects whose creator no longer exists (the reference leaks its
 * /dev/shm segments the same way: BytePS_ShM_* in src/rdma_van.h are never unlinked).
 */
#ifndef PS_VAN_SHM_UTIL_H_
#define PS_VAN_SHM_UTIL_H_
#include <dirent.h>
#include <signal.h>
#include <sys/mman.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>

namespace ps {

/*! \brief unlink /dev/shm/<prefix><pid>_* whose <pid> is gone; returns how many were removed */
inline int SweepStaleShm(const char* prefix) {
  DIR* d = opendir("/dev/shm");
  if (!d) return 0;
  const size_t plen = strlen(prefix);
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
