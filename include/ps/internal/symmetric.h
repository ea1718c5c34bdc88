/**
 * \file symmetric.h
 * \brief Symmetric memory: identical allocations in every process of a job, mapped by all of them
 *        (and, on NVSwitch, bound to one multicast address). See Van::AllocSymmetric.
 */
#ifndef PS_INTERNAL_SYMMETRIC_H_
#define PS_INTERNAL_SYMMETRIC_H_
#include <cstddef>
#include <string>
#include <utility>
#include <vector>

namespace ps {

/*!
 * \brief a block of memory that every process of a group allocated identically and mapped from
 *        every other member ("symmetric memory"): peers[i] is THIS process's mapping of member
 *        i's block (peers[index] == local). On an NVSwitch box `mc` is a multicast address
 *        bound to offset 0 of all blocks: one multimem.st there lands in every member's HBM,
 *        multimem.ld_reduce sums across them inside the switch (NVLS).
 */
struct SymmetricBuffer {
  void* local = nullptr;
  size_t bytes = 0;   // rounded up to the allocation granularity
  int index = 0;
  int count = 0;
  std::vector<void*> peers;
  void* mc = nullptr;
  /*! \brief (node id, member index) for every worker / server node of the job: which block is worker r's? */
  std::vector<std::pair<int, int>> node_member;
  int MemberOfNode(int node_id) const {
    for (const auto& nm : node_member) {
      if (nm.first == node_id) return nm.second;
    }
    return -1;
  }
};

/*! \brief who takes part in a symmetric allocation: the processes of a job on this host, in rank order */
struct SymmetricGroup {
  int job_port = 0;           // the scheduler's port: names the job on this host
  std::vector<int> pids;      // one entry per participating process
  int index = -1;             // this process's position in `pids`
};

}  // namespace ps
#endif  // PS_INTERNAL_SYMMETRIC_H_
