#!/usr/bin/env python3
"""Merge the PS_EVENT_TRACE files of the processes of one host into one timeline and print a
window of it (microseconds relative to the window start; one column per thread).
usage: scripts/event_timeline.py <file> [<file> ...] [--skip N] [--count N]"""
import sys


def main():
    files = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = {"--skip": -700, "--count": 400}
    args = sys.argv[1:]
    for i, a in enumerate(args):
        if a in opts:
            opts[a] = int(args[i + 1])
    files = [f for f in files if not f.lstrip("-").isdigit()]
    events = []
    for idx, path in enumerate(files):
        for line in open(path):
            ns, tid, tag, a, b = line.rstrip("\n").split("\t")
            events.append((int(ns), idx, int(tid), tag, int(a), int(b)))
    events.sort()
    skip = opts["--skip"]
    window = events[skip:skip + opts["--count"]] if skip >= 0 else events[skip:][:opts["--count"]]
    if not window:
        return
    threads = sorted({(e[1], e[2]) for e in window})
    col = {t: i for i, t in enumerate(threads)}
    print("# columns: " + "  ".join(f"[{i}]=file{t[0]}/tid{t[1]}" for t, i in col.items()))
    t0 = window[0][0]
    kinds = {0: "resp", 1: "resp(push)", 2: "pull", 3: "push"}
    for ns, idx, tid, tag, a, b in window:
        pad = "    " * col[(idx, tid)]
        print(f"{(ns - t0) / 1e3:9.1f}  {pad}[{col[(idx, tid)]}] {tag} ts={a} {kinds.get(b, b)}")


if __name__ == "__main__":
    main()
