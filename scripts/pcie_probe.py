#!/usr/bin/env python3
"""Host <-> device copy bandwidth of this box: H2D alone, D2H alone and both at once (pinned memory,
4,096,000 B chunks like bench.py's end-to-end pass, and one large copy). The end-to-end number of
bench.py can be at most min(H2D, D2H) when both directions overlap perfectly; this prints the bound.
Timing: CUDA events on the copy streams, after warm-up."""
import json
import sys

import torch


def timed(fn, streams, reps=5):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in streams]
        for (e0, _), s in zip(evs, streams):
            e0.record(s)
        fn()
        for (_, e1), s in zip(evs, streams):
            e1.record(s)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e1) for e0, e1 in evs))
    return best


def main():
    if not torch.cuda.is_available():
        print(json.dumps({"unavailable": "no CUDA device"}))
        return 0
    chunk, n = 4096000, 40
    host_a = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(n)]
    host_b = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(n)]
    dev_a = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(n)]
    dev_b = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(n)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s_in):
            for d, h in zip(dev_a, host_a):
                d.copy_(h, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s_out):
            for d, h in zip(dev_b, host_b):
                h.copy_(d, non_blocking=True)

    def both():
        h2d()
        d2h()

    total = chunk * n / 1e9
    out = {"chunk_bytes": chunk, "chunks": n,
           "h2d_GBps": total / (timed(h2d, [s_in]) * 1e-3),
           "d2h_GBps": total / (timed(d2h, [s_out]) * 1e-3)}
    ms = timed(both, [s_in, s_out])
    out["duplex_GBps_per_direction"] = total / (ms * 1e-3)
    big_h = torch.empty(chunk * n, dtype=torch.uint8).pin_memory()
    big_d = torch.empty(chunk * n, dtype=torch.uint8, device="cuda")

    def big():
        with torch.cuda.stream(s_in):
            big_d.copy_(big_h, non_blocking=True)
    out["h2d_single_copy_GBps"] = total / (timed(big, [s_in]) * 1e-3)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
