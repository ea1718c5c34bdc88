/**
 * \file van_factory.cc
 * \brief Transport selection. The RDMA-class names of the reference ("1",
 * "ibverbs", "ucx", "fabric"; src/van.cc:79-103, src/postoffice.cc:52-58) all
 * resolve to the one-sided NVLink van: on a B200 NVSwitch box peer HBM over
 * NVLink *is* the RDMA fabric.
 */
#include "van/van_factory.h"
#include "ps/internal/postoffice.h"
#include "van/tcp_van.h"

namespace ps {

Van* CreateVanByType(const std::string& type, Postoffice* postoffice) {
  LOG_IF(INFO, GetEnv("PS_VERBOSE", 0) >= 1) << "Creating Van: " << type;
  if (type == "zmq" || type == "0" || type == "tcp" || type.empty()) {
    return new TcpVan(postoffice);
  }
  LOG(FATAL) << "unsupported van type: " << type;
  return nullptr;
}

}  // namespace ps
