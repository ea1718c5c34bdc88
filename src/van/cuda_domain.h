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
 * cudaEvent (the default path is the copy engine: posted descriptors, completion
 * signalled by the kernel). A process drives one GPU (PS_CUDA_DEVICE, else LOCAL_RANK, else the
 * current device), several consecutive ones (DMLC_NUM_GPU_DEV: a context per device like
 * src/ucx_van.h:948-995), or one per instance of a group (PS_INSTANCE_GPU_STRIDE).
 */
#ifndef PS_VAN_CUDA_DOMAIN_H_
#define PS_VAN_CUDA_DOMAIN_H_
#include "van/mem_domain.h"

namespace ps {
/*! \brief nullptr (with a log line) if no CUDA device is usable */
MemDomain* CreateCudaDomain(int instance_idx = 0);
/*! \brief number of visible CUDA devices, 0 if the driver is absent */
int CudaDeviceCount();
}  // namespace ps
#endif  // PS_VAN_CUDA_DOMAIN_H_
