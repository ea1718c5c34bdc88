#!/usr/bin/env python
"""ResNet-50 synchronous PS training (BASELINE.json config 3), one process per GPU.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_resnet50_ps.py \
        --topology split --batch 128 --steps 50          # 4 workers + 4 servers
    python examples/train_resnet50_ps.py --steps 20      # 1 GPU, worker + server co-located

Parameters are kept in bf16 on the workers (channels-last convolutions under autocast-free
bf16), gradients are pushed in bf16 or block-scaled fp8 from autograd hooks, the servers
run fused SGD-momentum on fp32 master weights. Synthetic ImageNet-shaped data.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import pslite_b200  # noqa: E402
from pslite_b200.models.resnet import resnet50, resnet_tiny  # noqa: E402
from pslite_b200.parallel.launch import init_ps  # noqa: E402
from pslite_b200.parallel.ps_trainer import PSWorkerOptimizer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--topology", default="joint", choices=["joint", "split"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--grad-wire", default="bf16", choices=["bf16", "fp8"])
    ap.add_argument("--tiny", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("cpu:gloo,cuda:nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    C = pslite_b200.native()
    ctx = init_ps(args.topology)
    W, S = ctx.num_workers, ctx.num_servers
    server = None
    if ctx.is_server:
        server = C.GpuServer(0, num_workers=W, optimizer="sgd", lr=args.lr, beta1=0.9,
                             weight_decay=1e-4, grad_scale=1.0 / W, fuse_pull=True)
    if ctx.is_worker:
        model = (resnet_tiny() if args.tiny else resnet50()).to(dev).to(torch.bfloat16)
        model = model.to(memory_format=torch.channels_last)
        # (channels-last weights are dense: the PS moves each as the flat buffer of its storage)
        params = [p for p in model.parameters()]
        kv = C.KVWorker(0, 0)
        opt = PSWorkerOptimizer(params, kv, S, W, ctx.worker_rank, grad_wire=args.grad_wire,
                                chunk_elems=4 << 20).attach()
        opt.init_parameters(barrier=lambda: C.barrier(0, C.WORKER_GROUP, "worker"))
        hw = 32 if args.tiny else 224
        x = torch.randn(args.batch, 3, hw, hw, device=dev, dtype=torch.bfloat16)
        y = torch.randint(0, 10 if args.tiny else 1000, (args.batch,), device=dev)
        torch.cuda.synchronize()
        t0 = time.time()
        for step in range(args.steps):
            loss = F.cross_entropy(model(x).float(), y)
            loss.backward()
            opt.step()
            if step % 5 == 0 and ctx.worker_rank == 0:
                print(f"step {step} loss {loss.item():.4f}", flush=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        if ctx.worker_rank == 0:
            print(f"{args.batch * W * args.steps / dt:.1f} images/s over {W} worker(s), "
                  f"{opt.stats.keys} keys, final loss {loss.item():.4f}", flush=True)
    ctx.shutdown()


if __name__ == "__main__":
    main()
