"""Joint (worker + server engine in one process) PS training of a tiny model, compared with
a local fp32-master AdamW run; then a checkpoint round trip of the server state.
On a GPU: nvl van + device backend; without one: shm van + host backend, parameters in shared
memory. usage: train_joint.py <grad_wire> <steps>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pslite_b200  # noqa: E402
from pslite_b200.models.llama import Llama, LlamaConfig  # noqa: E402
from pslite_b200.parallel.launch import init_ps  # noqa: E402
from pslite_b200.parallel.ps_trainer import PSWorkerOptimizer  # noqa: E402


def main():
    wire, steps = sys.argv[1], int(sys.argv[2])
    C = pslite_b200.native()
    use_cuda = torch.cuda.is_available()
    dev = torch.device("cuda:0") if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(0)
    ctx = init_ps("joint", van="nvl" if use_cuda else "shm")
    lr, wd = 3e-3, 0.1
    server = C.GpuServer(0, num_workers=1, optimizer="adamw", lr=lr, beta1=0.9, beta2=0.95, eps=1e-8,
                         weight_decay=wd, grad_scale=1.0, fuse_pull=True)
    cfg = LlamaConfig.tiny()
    torch.manual_seed(0)
    with torch.device(dev):
        model = Llama(cfg).to(torch.bfloat16)
        ref = Llama(cfg).to(torch.bfloat16)
    model.init_weights(seed=1)
    ref.load_state_dict(model.state_dict())
    keep = []
    if not use_cuda:  # shared-memory parameters: pulls land in place like they do in HBM
        for p in model.parameters():
            buf = C.alloc_exportable(p.numel() * p.element_size(), "worker")
            view = buf.view(p.dtype).view(p.shape)
            view.copy_(p.data)
            p.data = view
            keep.append(buf)
    # local reference: fp32 master copy + torch AdamW, bf16 compute weights
    masters = [p.detach().float().clone().requires_grad_(True) for p in ref.parameters()]
    # same rule as PSWorkerOptimizer(no_decay_1d=True): norm gains are not decayed
    ropt = torch.optim.AdamW([{"params": [m for m in masters if m.dim() > 1], "weight_decay": wd},
                              {"params": [m for m in masters if m.dim() <= 1], "weight_decay": 0.0}],
                             lr=lr, betas=(0.9, 0.95), eps=1e-8)
    kv = C.KVWorker(0, 0)
    opt = PSWorkerOptimizer(model.parameters(), kv, 1, 1, 0, grad_wire=wire, chunk_elems=1 << 14).attach()
    opt.init_parameters(barrier=lambda: None)
    g = torch.Generator(device=dev).manual_seed(7)
    ps_losses, ref_losses = [], []
    tok = torch.randint(0, cfg.vocab_size, (2, 65), device=dev, generator=g)  # one fixed batch
    for _ in range(steps):
        loss = model(tok[:, :-1], tok[:, 1:])
        loss.backward()
        opt.step()
        ps_losses.append(loss.item())
        rl = ref(tok[:, :-1], tok[:, 1:])
        rl.backward()
        for mp, p in zip(masters, ref.parameters()):
            mp.grad = p.grad.float()
            p.grad = None
        ropt.step()
        with torch.no_grad():
            for mp, p in zip(masters, ref.parameters()):
                p.copy_(mp.to(torch.bfloat16))
        ref_losses.append(rl.item())
    if use_cuda:
        torch.cuda.synchronize()
    with torch.no_grad():
        worst = max(float((p.float() - q.float()).abs().max()) for p, q in zip(model.parameters(), ref.parameters()))
    print("PS  ", ["%.4f" % x for x in ps_losses])
    print("REF ", ["%.4f" % x for x in ref_losses])
    print(f"max_param_diff={worst:.5f} updates={server.num_updates()} fused={server.num_fused_fanouts()} "
          f"keys={server.num_keys()} launches={C.kernel_launch_count()}")
    ok = ps_losses[-1] < ps_losses[0] and abs(ps_losses[-1] - ref_losses[-1]) < (0.02 if wire == "bf16" else 0.08)
    ok = ok and server.num_fused_fanouts() > 0
    ok = ok and server.on_device() == use_cuda
    ok = ok and list(C.dead_nodes(60, "worker")) == []  # liveness query is callable from any node
    if not ok:
        print("FAIL (before the checkpoint stage)")
    # checkpoint round trip: save, train on, restore -> the fp32 master is back to the saved one
    # (validated on B200 in round 2: profiles/r2/01_verify_head_1gpu.log)
    import tempfile

    key = opt.chunks[0][0].key
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "server0.ckpt")
        ok = ok and server.save(path)
        saved = server.read_master(key).clone()
        for _ in range(2):
            model(tok[:, :-1], tok[:, 1:]).backward()
            opt.step()
        moved = not torch.equal(server.read_master(key), saved)
        ok = ok and moved and server.load(path)
        restored = server.read_master(key)
        ckpt_ok = torch.equal(restored, saved)
        print(f"checkpoint: moved_after_save={moved} restored_equal={ckpt_ok}")
        ok = ok and ckpt_ok
    ctx.shutdown()
    print("PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
