/**
 * \file cuda_domain.h
 * \brief CudaDomain: B200 HBM as peer-mappable memory (the "nvl" van backend).
 *
 * Export = cuMemGetAddressRange + cudaIpcGetMemHandle on the containing
 * allocation (works for cudaMalloc memory, which includes PyTorch's caching
 * allocator segments); Import = cudaIpcOpenMemHandle with lazy peer access, or
 * the raw pointer when exporter and importer share a process (co-located
 * worker + server: the reference's IPCTransport case, src/rdma_van.h:231-244).
 * CopyAsync launches the sm_100a copy / cast / fp8-quant kernels of
 * src/kernels on a dedicated high-priority stream; the Ticket is a pooled
 * cudaEvent. One process drives one GPU (UCXVan picks a context per device,
 * src/ucx_van.h:948-995; here: PS_CUDA_DEVICE, else LOCAL_RANK, else the
 * current device).
 */
#ifndef PS_VAN_CUDA_DOMAIN_H_
#define PS_VAN_CUDA_DOMAIN_H_
#include "van/mem_domain.h"

namespace ps {
/*! \brief nullptr (with a log line) if no CUDA device is usable */
MemDomain* CreateCudaDomain();
/*! \brief number of visible CUDA devices, 0 if the driver is absent */
int CudaDeviceCount();
}  // namespace ps
#endif  // PS_VAN_CUDA_DOMAIN_H_
