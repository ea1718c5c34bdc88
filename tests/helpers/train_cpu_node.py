"""CPU end-to-end check of the PS trainer logic (hooks, chunking, init, sync rounds).

usage: train_cpu_node.py <role> <num_workers> <port> <steps>
A Python KVServer implements synchronous SGD on fp32 masters; workers train a tiny bf16
MLP through PSWorkerOptimizer over the TCP van. Every worker prints a parameter checksum;
the test compares them with each other and with a local simulation.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pslite_b200  # noqa: E402
from pslite_b200.parallel.ps_trainer import PSWorkerOptimizer  # noqa: E402

LR = 0.05


def make_model():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    return m.to(torch.bfloat16)


def batch(rank: int, step: int):
    g = torch.Generator().manual_seed(1000 * rank + step)
    x = torch.randn(8, 16, generator=g).to(torch.bfloat16)
    y = torch.randn(8, 4, generator=g).to(torch.bfloat16)
    return x, y


def simulate(nw: int, steps: int) -> float:
    """what the PS job must compute: average the workers' bf16 grads, SGD on fp32 masters"""
    model = make_model()
    masters = [p.detach().float().clone() for p in model.parameters()]
    for s in range(steps):
        acc = [torch.zeros_like(m) for m in masters]
        for r in range(nw):
            x, y = batch(r, s)
            model.zero_grad(set_to_none=True)
            torch.nn.functional.mse_loss(model(x).float(), y.float()).backward()
            for a, p in zip(acc, model.parameters()):
                a += p.grad.float()
        with torch.no_grad():
            for m, a, p in zip(masters, acc, model.parameters()):
                m -= LR * a / nw
                p.copy_(m.to(torch.bfloat16))
    return float(sum(p.double().sum() for p in model.parameters()))


def main():
    role, nw, port, steps = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    C = pslite_b200.native()
    for k, v in {"DMLC_NUM_WORKER": nw, "DMLC_NUM_SERVER": 1, "DMLC_PS_ROOT_URI": "127.0.0.1",
                 "DMLC_PS_ROOT_PORT": port, "DMLC_NODE_HOST": "127.0.0.1", "PS_VAN_TYPE": "zmq",
                 "DMLC_ROLE": role}.items():
        C.set_env(k, str(v))
    C.start_ps(0, role, -1, True)
    if role == "scheduler":
        C.finalize(0, role, True)
        return
    if role == "server":
        srv = C.KVServer(0)
        state = {}  # key -> dict(master, acc, pushed:set, waiting:list)

        def handle(meta, key, vals):
            st = state.setdefault(key, {"master": None, "acc": None, "pushed": set(), "waiting": []})
            rank = meta["sender_rank"]
            if meta["push"] and meta["cmd"] == C.CMD_INIT_BF16:
                if st["master"] is None:
                    st["master"] = vals.view(torch.bfloat16).float().clone()
                srv.response(meta["id"])
            elif meta["push"]:
                g = vals.view(torch.bfloat16).float()
                st["acc"] = g.clone() if st["acc"] is None else st["acc"] + g
                st["pushed"].add(rank)
                srv.response(meta["id"])
                if len(st["pushed"]) == nw:
                    st["master"] -= LR * st["acc"] / nw
                    st["acc"] = None
                    st["pushed"].clear()
                    out = st["master"].to(torch.bfloat16).view(torch.uint8)
                    for rid in st["waiting"]:
                        srv.response(rid, out)
                    st["waiting"].clear()
            else:
                if rank in st["pushed"]:
                    st["waiting"].append(meta["id"])
                else:
                    srv.response(meta["id"], st["master"].to(torch.bfloat16).view(torch.uint8))

        srv.set_request_handle(handle)
        C.finalize(0, role, True)
        return
    rank = C.worker_rank()
    model = make_model()
    kv = C.KVWorker(0, 0)
    opt = PSWorkerOptimizer(model.parameters(), kv, 1, nw, rank, grad_wire="bf16", chunk_elems=1024).attach()
    opt.init_parameters(barrier=lambda: C.barrier(0, C.WORKER_GROUP, "worker"))
    assert opt.stats.keys >= 4
    for s in range(steps):
        x, y = batch(rank, s)
        torch.nn.functional.mse_loss(model(x).float(), y.float()).backward()
        opt.step()
    chk = float(sum(p.double().sum() for p in model.parameters()))
    print(f"CHECKSUM {rank} {chk:.6f} EXPECT {simulate(nw, steps):.6f}", flush=True)
    C.finalize(0, role, True)


if __name__ == "__main__":
    main()
