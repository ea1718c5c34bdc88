"""One node of a tiny Python-level PS job (used by the CPU tests).

usage: ps_node.py <role> <van> <num_workers> <num_servers> <port> [n_elems]
The server sums float32 pushes per key in Python; workers push, pull, and check.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pslite_b200  # noqa: E402


def main():
    role, van, nw, ns, port = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    n = int(sys.argv[6]) if len(sys.argv) > 6 else 1000
    C = pslite_b200.native()
    for k, v in {"DMLC_NUM_WORKER": nw, "DMLC_NUM_SERVER": ns, "DMLC_PS_ROOT_URI": "127.0.0.1",
                 "DMLC_PS_ROOT_PORT": port, "DMLC_NODE_HOST": "127.0.0.1", "PS_VAN_TYPE": van,
                 "DMLC_ROLE": role}.items():
        C.set_env(k, str(v))
    C.start_ps(0, role, -1, True)
    if role == "server":
        store = {}
        srv = C.KVServer(0)

        def handle(meta, key, vals):
            if meta["push"]:
                x = vals.view(torch.float32).clone()
                store[key] = store[key] + x if key in store else x
                srv.response(meta["id"])
            else:
                srv.response(meta["id"], store[key].view(torch.uint8))

        srv.set_request_handle(handle)
        # control messages next to the KV traffic: reply with the reversed body
        ctl = C.SimpleApp(7, 7)  # on servers a customer is found by its app id (as in the reference)
        ctl.set_request_handle(lambda head, body, sender, ts: body[::-1] if head == 3 else None)
        C.finalize(0, role, True)
        return
    if role == "worker":
        kv = C.KVWorker(0, 0)
        rank = C.worker_rank()
        one_sided = van == "shm"
        ok = True
        for i in range(ns * 2):
            key = kv.server_key(i % ns, i)
            if one_sided:
                buf = C.alloc_exportable(n * 4).view(torch.float32)
                out = C.alloc_exportable(n * 4).view(torch.float32)
            else:
                buf = torch.empty(n)
                out = torch.empty(n)
            buf.copy_(torch.arange(n, dtype=torch.float32) * (i + 1))
            kv.wait(kv.push(key, buf.view(torch.uint8)))
            C.barrier(0, C.WORKER_GROUP, "worker")       # every worker has pushed key i
            out.zero_()
            kv.wait(kv.pull(key, out.view(torch.uint8)))
            expect = torch.arange(n, dtype=torch.float32) * (i + 1) * nw
            if not torch.allclose(out, expect):
                ok = False
                print(f"worker {rank}: key {i} mismatch {out[:4]} vs {expect[:4]}", flush=True)
            C.barrier(0, C.WORKER_GROUP, "worker")
        # SimpleApp: one request to every server, each answers with the reversed body
        ctl = C.SimpleApp(7, 1)
        answers = []
        ctl.set_response_handle(lambda head, body, sender, ts: answers.append((head, body, sender)))
        ctl.wait(ctl.request(3, f"hello from {rank}", C.SERVER_GROUP))
        want = f"hello from {rank}"[::-1].encode()
        if len(answers) != ns or any(h != 3 or b != want for h, b, _ in answers):
            ok = False
            print(f"worker {rank}: SimpleApp answers {answers}", flush=True)
        ctl.wait(ctl.request(4, "", C.SERVER_GROUP))  # a handler that returns None: empty reply
        if len(answers) != 2 * ns or answers[-1][1] != b"":
            ok = False
            print(f"worker {rank}: SimpleApp empty answers {answers}", flush=True)
        print(f"worker {rank}: {'PASS' if ok else 'FAIL'}", flush=True)
        C.finalize(0, role, True)
        sys.exit(0 if ok else 1)
    C.finalize(0, role, True)  # scheduler


if __name__ == "__main__":
    main()
