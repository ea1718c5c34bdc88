"""bench.py's push/pull loop on host memory (shm or tcp van) under torchrun: the same
call sequence (init_ps, BenchServer, per-key warm-up push, push_pull_batch + wait_all rounds,
shutdown) with W workers and S servers, so that N-process hangs reproduce without a GPU.
usage: torchrun --nproc-per-node N pushpull_multi.py <van> <len> <keys_per_server> <rounds> [topology]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pslite_b200  # noqa: E402
from pslite_b200.parallel.launch import init_ps  # noqa: E402


def main():
    van, length, kps, rounds = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    topo = sys.argv[5] if len(sys.argv) > 5 else "split"
    fused = os.environ.get("PSLITE_TEST_PUSHPULL", "0") == "1"  # one ZPushPull per key instead of ZPush + ZPull
    staged = os.environ.get("PSLITE_TEST_STAGED", "0") == "1"  # KVWorker.staged_push_pull rounds
    rank = int(os.environ["RANK"])
    dist.init_process_group("gloo")
    C = pslite_b200.native()
    ctx = init_ps(topo, van=van)
    server = C.BenchServer(0) if ctx.is_server else None  # noqa: F841
    S, W = ctx.num_servers, ctx.num_workers
    total = S * kps
    t0 = time.time()
    if ctx.is_worker:
        kv = C.KVWorker(0, 0)
        keys = [kv.server_key(k % S, k) for k in range(total)]
        if van == "shm":  # one-sided path: the values must live in exportable (shared) memory
            vals = [C.alloc_exportable(length, "worker").fill_(1 + ctx.worker_rank) for _ in range(total)]
        else:
            vals = [torch.full((length,), 1 + ctx.worker_rank, dtype=torch.uint8) for _ in range(total)]
        for k in range(total):
            kv.wait(kv.push(keys[k], vals[k], order_after_current_stream=False))
    dist.barrier()
    if ctx.is_worker:
        host_in = [torch.full((length,), 1 + ctx.worker_rank, dtype=torch.uint8) for _ in range(total)] if staged else None
        host_out = [torch.zeros(length, dtype=torch.uint8) for _ in range(total)] if staged else None
        for _ in range(rounds):
            if staged:
                for h in host_out:
                    h.zero_()
                kv.staged_push_pull(keys, vals, host_in, host_out)
                for k in range(total):  # the pulled bytes reached host memory before the call returned
                    assert torch.equal(host_out[k], vals[k]), f"key {k}: host copy differs from the pulled tensor"
            elif fused:
                ts = [kv.push_pull(keys[k], vals[k], vals[k], order_after_current_stream=False) for k in range(total)]
                for t in ts:
                    kv.wait(t)
            else:
                kv.wait_all(kv.push_pull_batch(keys, vals, order_after_current_stream=False))
        # every pull must have returned the value stored by the key's first pusher
        for k in range(total):
            lo, hi = int(vals[k].min()), int(vals[k].max())
            assert lo == hi and 1 <= lo <= W, f"key {k}: pulled bytes in [{lo}, {hi}]"
    dist.barrier()
    dt = time.time() - t0
    ctx.shutdown()
    dist.destroy_process_group()
    if rank == 0:
        print(f"PASS {W}w+{S}s van={van} len={length} keys={total} rounds={rounds} {dt:.2f}s", flush=True)


if __name__ == "__main__":
    main()
