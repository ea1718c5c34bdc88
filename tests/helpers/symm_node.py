"""One process of a symmetric-memory check (CPU: shm van; GPU: nvl van under torchrun).

usage: symm_node.py <role> <van> <num_workers> <num_servers> <port>      (role: scheduler/server/worker)
Every worker and server process allocates the same symmetric buffer through the native runtime
(C.alloc_symmetric: FdExchange + memfd on the shm van, cuMemCreate / cuMulticast on the nvl van),
writes its own pattern, and after a barrier reads every other member's block through its mapping.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pslite_b200  # noqa: E402


def main():
    role, van, nw, ns, port = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    C = pslite_b200.native()
    for k, v in {"DMLC_NUM_WORKER": nw, "DMLC_NUM_SERVER": ns, "DMLC_PS_ROOT_URI": "127.0.0.1",
                 "DMLC_PS_ROOT_PORT": port, "DMLC_NODE_HOST": "127.0.0.1", "PS_VAN_TYPE": van,
                 "DMLC_ROLE": role}.items():
        C.set_env(k, str(v))
    dev = -1
    if van == "nvl" and role != "scheduler":
        dev = int(os.environ.get("PS_CUDA_DEVICE", "0"))
        torch.cuda.set_device(dev)
    C.start_ps(0, role, -1, True)
    if role == "scheduler":
        C.finalize(0, role, True)
        return
    n = 1 << 16
    local, mc, peers, index, count, wblocks, sblocks = C.alloc_symmetric("check", n, role)
    assert len(wblocks) == nw and len(sblocks) == ns and all(wblocks) and all(sblocks)
    mine = (wblocks if role == "worker" else sblocks)
    assert local.data_ptr() in mine, "this process's block must be listed under its own rank"
    assert count == nw + ns and len(peers) == count and local.numel() >= n
    assert int(local.sum()) == 0, "a fresh symmetric block is zero-filled"
    local[:n] = (torch.arange(n, dtype=torch.int32, device=local.device) * (index + 1) % 251).to(torch.uint8)
    if dev >= 0:
        torch.cuda.synchronize()
    # second call (any role of this process) returns the same block
    again = C.alloc_symmetric("check", n, role)
    assert again[0].data_ptr() == local.data_ptr() and again[3] == index
    C.barrier(0, C.WORKER_GROUP + C.SERVER_GROUP, role)
    ok = True
    for i, addr in enumerate(peers):
        got = C.read_bytes(addr, n, dev >= 0)  # on a GPU: read over NVLink through the peer mapping
        want = (torch.arange(n, dtype=torch.int32) * (i + 1) % 251).to(torch.uint8)
        ok = ok and torch.equal(got, want)
    C.barrier(0, C.WORKER_GROUP + C.SERVER_GROUP, role)
    if mc and dev >= 0:
        # NVLS: member 0 stores ONE stream to the multicast address, every member finds it in its own block
        if index == 0:
            src = (torch.arange(n, dtype=torch.int32, device=local.device) % 199).to(torch.uint8)
            C.copy_multicast(mc, src)
            torch.cuda.synchronize()
        C.barrier(0, C.WORKER_GROUP + C.SERVER_GROUP, role)
        want = (torch.arange(n, dtype=torch.int32) % 199).to(torch.uint8)
        ok = ok and torch.equal(local[:n].cpu(), want)
        C.barrier(0, C.WORKER_GROUP + C.SERVER_GROUP, role)
    print(f"{role} member {index}/{count} multicast={'yes' if mc else 'no'} {'PASS' if ok else 'FAIL'}", flush=True)
    C.finalize(0, role, True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
