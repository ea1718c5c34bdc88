/**
 * \file nvtx.h
 * \brief Optional NVTX ranges around the data-plane steps (PS_NVTX=1), for ncu / nsys timelines.
 *
 * The reference's only tracing is the van's text log (ENABLE_PROFILING, kept here with the same
 * format); on a GPU box the natural companion is NVTX: push / pull-reply / update / descriptor
 * ranges line up with the kernels in the profiler. nvtx3 is header-only and binds to the tools'
 * injection library lazily, so there is no link dependency and no cost when no tool is attached.
 */
#ifndef PS_CORE_NVTX_H_
#define PS_CORE_NVTX_H_
#include "ps/internal/utils.h"

#ifdef PS_USE_CUDA
#include <nvtx3/nvToolsExt.h>
#endif

namespace ps {

inline bool NvtxEnabled() {
  static const bool on = GetEnv("PS_NVTX", 0) != 0;
  return on;
}

/*! \brief RAII range; a no-op unless PS_NVTX=1 in a CUDA build */
class NvtxRange {
 public:
  explicit NvtxRange(const char* name) {
#ifdef PS_USE_CUDA
    if (NvtxEnabled()) {
      nvtxRangePushA(name);
      active_ = true;
    }
#else
    (void)name;
#endif
  }
  ~NvtxRange() {
#ifdef PS_USE_CUDA
    if (active_) nvtxRangePop();
#endif
  }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;

 private:
  bool active_ = false;
};

}  // namespace ps
#endif  // PS_CORE_NVTX_H_
