"""Multi-GPU PS training under torchrun (one process per GPU). Every rank trains the same
tiny Llama on its own data shard through the PS; at the end all workers must hold
bit-identical parameters (they all pulled the same server state) and the loss must drop.
usage: torchrun ... train_multi.py <topology> <grad_wire> <steps> [symm|nvls|async]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pslite_b200  # noqa: E402
from pslite_b200.models.llama import Llama, LlamaConfig  # noqa: E402
from pslite_b200.parallel.launch import init_ps  # noqa: E402
from pslite_b200.parallel.ps_trainer import (PSWorkerOptimizer, setup_symmetric_grads,  # noqa: E402
                                              setup_symmetric_params, symmetric_layout)


def main():
    topo, wire, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    symmetric = len(sys.argv) > 4 and sys.argv[4] in ("symm", "nvls")
    nvls_reduce = len(sys.argv) > 4 and sys.argv[4] == "nvls"
    # asynchronous SGD: every push is its own optimizer step, workers never wait for each other
    async_sgd = len(sys.argv) > 4 and sys.argv[4] == "async"
    fused_pp = os.environ.get("PSLITE_TEST_PUSHPULL", "0") == "1"  # KVWorker::ZPushPull per chunk
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    use_cuda = torch.cuda.is_available()
    dist.init_process_group("cpu:gloo,cuda:nccl" if use_cuda else "gloo")
    gloo = dist.new_group(backend="gloo")
    C = pslite_b200.native()
    if use_cuda:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    # without a GPU the same job runs on the CPU engine: shm one-sided van + host-resident shards
    ctx = init_ps(topo, van="nvl" if use_cuda else "shm")
    exportable = os.environ.get("PSLITE_TEST_EXPORTABLE_PARAMS", "0") == "1"
    W, S = ctx.num_workers, ctx.num_servers
    server = None
    if ctx.is_server:
        server = C.GpuServer(0, num_workers=W, optimizer="adamw", lr=3e-3 / (W if async_sgd else 1), beta1=0.9,
                             beta2=0.95, eps=1e-8, weight_decay=0.0, grad_scale=1.0 if async_sgd else 1.0 / W,
                             fuse_pull=True, async_updates=async_sgd)
    # checkpoint / resume: every server keeps its shards in <dir>/server<rank>.ckpt
    ckpt_dir = os.environ.get("PSLITE_CKPT_DIR", "")
    ckpt = os.path.join(ckpt_dir, f"server{ctx.server_rank}.ckpt") if ckpt_dir and server is not None else ""
    resumed = False
    if ckpt and os.path.exists(ckpt):
        resumed = server.load(ckpt)  # before any worker initialises: loaded shards win over init pushes
        assert resumed, f"cannot load {ckpt}"
    dist.barrier(group=gloo)
    ok = True
    checksum = torch.zeros(1, dtype=torch.float64)
    losses = []
    cfg = LlamaConfig.tiny()
    model = None
    if ctx.is_worker:
        with torch.device(dev):
            model = Llama(cfg).to(torch.bfloat16)
        model.init_weights(seed=3)
        if exportable and not use_cuda:
            # parameters in shared memory: pulls land in place (zero-copy), as they do in HBM
            keep = []
            for p in model.parameters():
                buf = C.alloc_exportable(p.numel() * p.element_size(), "worker")
                view = buf.view(p.dtype).view(p.shape)
                view.copy_(p.data)
                p.data = view
                keep.append(buf)
    mcast_info = ""
    if symmetric:
        # every rank (servers too) allocates the same symmetric buffer; workers move their
        # parameters into it; servers learn the multicast + peer addresses
        with torch.device("meta"):
            shapes = [p for p in Llama(cfg).parameters()]
        _, total = symmetric_layout(shapes)
        plist = list(model.parameters()) if model is not None else None
        flat, hdl, mc, peers, nbytes = setup_symmetric_params(plist, total, dist.group.WORLD, dev,
                                                              list(range(W)))
        mcast_info = f" multicast_ptr={'yes' if mc else 'NO'}"
        if server is not None:
            server.set_symmetric(mc, peers, nbytes)
    gbuf = None
    if nvls_reduce:
        gbuf, ghdl, gmc, gbytes = setup_symmetric_grads(total, dist.group.WORLD, dev,
                                                           "worker" if ctx.is_worker else "server")
        if not gmc:  # no NVSwitch multicast here: fall back to landing slots
            nvls_reduce = False
            mcast_info += " nvls=unavailable"
        elif server is not None:
            server.set_symmetric_grads(gmc, gbytes)
        dist.barrier(group=gloo)
    if ctx.is_worker:
        kv = C.KVWorker(0, 0)
        opt = PSWorkerOptimizer(model.parameters(), kv, S, W, ctx.worker_rank, grad_wire=wire,
                                chunk_elems=1 << 14, symmetric=symmetric,
                                grad_buffer=gbuf if nvls_reduce else None, fused_pushpull=fused_pp).attach()
        opt.init_parameters(barrier=lambda: C.barrier(0, C.WORKER_GROUP, "worker"))
        if os.environ.get("PSLITE_TEST_LAZY", "0") == "1":
            opt.enable_lazy_wait(model)  # step() returns at once; modules wait for their own parameters
        g = torch.Generator(device=dev).manual_seed(100 + ctx.worker_rank)
        tok = torch.randint(0, cfg.vocab_size, (2, 65), device=dev, generator=g)
        freeze_after = int(os.environ.get("PSLITE_TEST_FREEZE_AFTER", "0"))  # learning rate -> 0 from there
        ckpt_at = int(os.environ.get("PSLITE_CKPT_AT", "0"))  # joint topology: also save after that many steps
        for it in range(steps):
            if ckpt_at and it == ckpt_at and ckpt:
                opt.wait_all()
                dist.barrier(group=gloo)   # every worker has finished step `ckpt_at`
                assert server.save(ckpt)
                dist.barrier(group=gloo)
            if freeze_after and it == freeze_after:
                opt.wait_all()
                opt.set_lr(0.0)
                C.barrier(0, C.WORKER_GROUP, "worker")  # every server has the new rate before anyone pushes
            loss = model(tok[:, :-1], tok[:, 1:])
            loss.backward()
            opt.step()
            losses.append(loss.item())
        opt.wait_all()
        if use_cuda:
            torch.cuda.synchronize()
        with torch.no_grad():
            checksum[0] = sum(float(p.double().sum()) for p in model.parameters())
        ok = losses[-1] < losses[0]
        if freeze_after or ckpt:
            print(f"rank {rank}: all losses {['%.4f' % x for x in losses]}", flush=True)
    if ckpt and not int(os.environ.get("PSLITE_CKPT_AT", "0")):  # (a mid-run save is kept as it is)
        dist.barrier(group=gloo)  # every worker has finished its last step
        assert server.save(ckpt)
        print(f"rank {rank}: checkpoint {'resumed and ' if resumed else ''}saved to {ckpt}", flush=True)
    # all workers pulled the same parameters
    sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(sums, checksum, group=gloo)
    wsums = [float(s) for i, s in enumerate(sums) if (topo == "joint" or i < W)]
    # (asynchronous workers pull at different moments: their copies legitimately differ)
    same = async_sgd or all(abs(x - wsums[0]) < 1e-9 for x in wsums)
    engine = ("device" if server.on_device() else "host") if server else "-"
    print(f"rank {rank}: engine={engine} losses {['%.3f' % l for l in losses[:2]]}..{['%.3f' % l for l in losses[-2:]]} "
          f"checksums_equal={same} updates={server.num_updates() if server else 0} "
          f"fused={server.num_fused_fanouts() if server else 0} "
          f"mcast={server.num_multicast_fanouts() if server else 0} "
          f"switch_reduce={server.num_switch_reductions() if server else 0}{mcast_info}", flush=True)
    if nvls_reduce and server is not None:
        ok = ok and server.num_switch_reductions() == server.num_updates() > 0
    ok = ok and same
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=gloo)
    ctx.shutdown()
    dist.destroy_process_group()
    if rank == 0:
        print("PASS" if flag.item() > 0 else "FAIL", flush=True)
    sys.exit(0 if flag.item() > 0 else 1)


if __name__ == "__main__":
    main()
