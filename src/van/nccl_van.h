/**
 * \file nccl_van.h
 * \brief NcclVan: two-sided GPU transport — payloads travel as ncclSend / ncclRecv
 *        pairs, descriptors over the inherited TCP / shared-memory channel.
 *
 * Counterpart of the reference's two-sided, rendezvous-based FabricVan
 * (src/fabric_van.h:50-1173, src/fabric_transport.h: tagged send/recv whose
 * endpoints are bootstrapped through an inner ZMQ van) for a CUDA cluster:
 * the "fabric" is whatever NCCL finds between two GPUs — NVLink / NVSwitch inside
 * an NVL domain, GPUDirect RDMA between nodes. It is the van to use when peers
 * cannot map each other's memory (different OS instances, containers without a
 * shared IPC namespace), i.e. where the one-sided `nvl` van does not apply.
 *
 *   endpoint bootstrap   fi_av_insert of an opaque address blob carried by the inner
 *                        van  ->  a ncclUniqueId carried by an ADDR_REQUEST control
 *                        message; each *directed* pair (src -> dst) gets its own
 *                        2-rank communicator, created on first use
 *   tagged send / recv   the descriptor (Meta) is sent first on the TCP channel, in the
 *                        same critical section that enqueues ncclSend on the pair's
 *                        stream; the receiving van thread reads the descriptor, picks
 *                        the destination (registered buffer | the pull's own tensor |
 *                        a cached landing buffer) and enqueues the matching ncclRecv.
 *                        One direction per communicator => send order == recv order,
 *                        no cross-dependency between the two directions, no deadlock.
 *   completion           a delivery thread waits for the recv's event and only then
 *                        hands the message (payload already in HBM) to the customer;
 *                        later messages queue behind it, so per-sender order holds
 *   wire transforms      the fused cast / fp8 block-quant kernels run on the pair's
 *                        stream into a staging buffer in front of ncclSend
 * Peers that share this process's GPU (co-located worker + server) cannot form a NCCL
 * communicator; for them — and for GPU-less peers — the payload is staged through host
 * memory on the TCP channel and placed back on the device by the receiver.
 */
#ifndef PS_VAN_NCCL_VAN_H_
#define PS_VAN_NCCL_VAN_H_

namespace ps {
class Postoffice;
class Van;
/*! \brief nullptr (with a log line) when CUDA or libnccl is not usable in this process */
Van* CreateNcclVan(Postoffice* postoffice);
}  // namespace ps
#endif  // PS_VAN_NCCL_VAN_H_
