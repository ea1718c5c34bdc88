#!/usr/bin/env python3
"""Turn the raw stacks written by PS_SAMPLE_PROFILE (src/core/sampler.h) into a flat profile:
self time by function and inclusive time by function, using addr2line on the executable.
usage: scripts/symbolize_samples.py <binary> <samples file> [top N]"""
import collections
import re
import subprocess
import sys


def main():
    binary, path = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 35
    base = None
    stacks = []
    tids = []
    for line in open(path):
        if line.startswith("#map"):
            m = re.match(r"#map ([0-9a-f]+)-[0-9a-f]+ r[-w]-p 00000000 \S+ \S+\s+(\S+)", line)
            if m and base is None and m.group(2).endswith(binary.split("/")[-1]):
                base = int(m.group(1), 16)
            continue
        parts = line.split()
        tid = parts[0] if parts and parts[0].startswith("t") else "t?"
        addrs = [int(a, 16) for a in parts if not a.startswith("t")]
        if addrs:
            stacks.append(addrs)
            tids.append(tid)
    base = base or 0
    uniq = sorted({a for s in stacks for a in s})
    rel = [hex(a - base) if a >= base else hex(a) for a in uniq]
    out = subprocess.run(["addr2line", "-f", "-C", "-i", "-e", binary] + rel, capture_output=True, text=True).stdout
    # without -i parsing complexity: re-run without inlines for a 2-lines-per-address layout
    out = subprocess.run(["addr2line", "-f", "-C", "-e", binary] + rel, capture_output=True, text=True).stdout.splitlines()
    name = {}
    for i, a in enumerate(uniq):
        fn = out[2 * i] if 2 * i < len(out) else "??"
        name[a] = re.sub(r"\(.*", "", fn)[:90]
    self_t, incl_t = collections.Counter(), collections.Counter()
    for s in stacks:
        self_t[name[s[0]]] += 1
        for fn in {name[a] for a in s}:
            incl_t[fn] += 1
    n = len(stacks)
    per_thread = collections.Counter(tids)
    print("-- CPU by thread: " + ", ".join(f"{t}={c}" for t, c in per_thread.most_common(8)))
    busiest = per_thread.most_common(3)
    for t, c in busiest:
        incl = collections.Counter()
        for s, tt in zip(stacks, tids):
            if tt == t:
                for fn in {name[a] for a in s}:
                    incl[fn] += 1
        print(f"-- thread {t} ({c} samples), inclusive:")
        for fn, k in incl.most_common(14):
            print(f"   {100.0 * k / c:6.2f}%  {fn}")
    print(f"{n} samples (1 ms of CPU each)\n-- self")
    for fn, c in self_t.most_common(top):
        print(f"{100.0 * c / n:6.2f}%  {fn}")
    print("-- inclusive")
    for fn, c in incl_t.most_common(top):
        print(f"{100.0 * c / n:6.2f}%  {fn}")


if __name__ == "__main__":
    main()
