/**
 * \file van_factory.cc
 * \brief Transport selection. The RDMA-class names of the reference ("1",
 * "ibverbs", "ucx", "fabric"; src/van.cc:79-103, src/postoffice.cc:52-58) all
 * resolve to the one-sided NVLink van: on a B200 NVSwitch box peer HBM over
 * NVLink *is* the RDMA fabric. "nccl" selects the two-sided NcclVan (the FabricVan
 * counterpart) for peers that cannot map each other's memory.
 */
#include "van/van_factory.h"
#include "ps/internal/postoffice.h"
#include "van/multi_van.h"
#include "van/onesided_van.h"
#include "van/tcp_van.h"
#ifdef PS_USE_CUDA
#include "van/cuda_domain.h"
#include "van/nccl_van.h"
#endif

namespace ps {

Van* CreateVanByType(const std::string& type, Postoffice* postoffice) {
  LOG_IF(INFO, GetEnv("PS_VERBOSE", 0) >= 1) << "Creating Van: " << type;
  if (type == "zmq" || type == "0" || type == "tcp" || type.empty()) {
    return new TcpVan(postoffice);
  }
  if (type == "multivan") return new MultiVan(postoffice);
  if (type == "shm") return new OneSidedVan(postoffice, new ShmDomain(), "shm");
  if (type == "nvl" || type == "1" || type == "ibverbs" || type == "ucx" || type == "fabric") {
#ifdef PS_USE_CUDA
    MemDomain* dom = CreateCudaDomain(postoffice ? postoffice->instance_idx() : 0);
    CHECK(dom) << "van type '" << type << "' maps to the NVLink van, which needs a GPU";
    return new OneSidedVan(postoffice, dom, "nvl");
#else
    LOG(FATAL) << "van type '" << type << "' needs a build with PS_USE_CUDA";
#endif
  }
  if (type == "nccl") {
#ifdef PS_USE_CUDA
    Van* van = CreateNcclVan(postoffice);
    CHECK(van) << "van type 'nccl' needs a GPU and libnccl.so.2";
    return van;
#else
    LOG(FATAL) << "van type 'nccl' needs a build with PS_USE_CUDA";
#endif
  }
  LOG(FATAL) << "unsupported van type: " << type;
  return nullptr;
}

}  // namespace ps
