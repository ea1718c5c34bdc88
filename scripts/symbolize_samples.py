#!/usr/bin/env python3
"""Turn the raw stacks written by PS_SAMPLE_PROFILE (src/core/sampler.h) into flat profiles, one per
busy thread: self time (the innermost frame that is not the signal machinery) and inclusive time by
function. Addresses inside the executable are resolved with addr2line, addresses inside shared
libraries (libc, libstdc++: stripped, but their dynamic symbols are there) with `nm -D`.
usage: scripts/symbolize_samples.py <binary> <samples file> [top N]"""
import bisect
import collections
import re
import subprocess
import sys

SIGNAL_FRAMES = ("ps::SampleProfiler", "__restore_rt", "backtrace", "_Unwind", "__sigaction", "killpg", "??")


def load(path):
    maps, stacks, tids = [], [], []
    seen = set()
    for line in open(path):
        if line.startswith("#map"):
            m = re.match(r"#map ([0-9a-f]+)-([0-9a-f]+) (\S+) ([0-9a-f]+) \S+ \S+\s+(\S+)", line)
            if m and "x" in m.group(3) and m.group(5).startswith("/"):
                key = (m.group(1), m.group(5))
                if key not in seen:
                    seen.add(key)
                    maps.append((int(m.group(1), 16), int(m.group(2), 16), int(m.group(4), 16), m.group(5)))
            continue
        parts = line.split()
        tid = parts[0] if parts and parts[0].startswith("t") else "t?"
        addrs = []
        for a in parts:
            if a.startswith("t"):
                continue
            try:
                addrs.append(int(a, 16))
            except ValueError:
                pass
        if addrs:
            stacks.append(addrs)
            tids.append(tid)
    return maps, stacks, tids


def dynsyms(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", lib], capture_output=True, text=True).stdout
    table = []
    for line in out.splitlines():
        p = line.split(None, 2)
        if len(p) == 3 and p[1] in "TtWwiI":
            table.append((int(p[0], 16), re.sub(r"@.*", "", p[2])))
    table.sort()
    return [a for a, _ in table], [n for _, n in table]


def main():
    binary, path = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    maps, stacks, tids = load(path)
    uniq = sorted({a for s in stacks for a in s})
    name = {}
    by_file = collections.defaultdict(list)
    for a in uniq:
        for lo, hi, off, f in maps:
            if lo <= a < hi:
                by_file[f].append((a, a - lo + off))
                break
        else:
            name[a] = "??"
    for f, items in by_file.items():
        if f.endswith(binary.split("/")[-1]):
            rel = [hex(r) for _, r in items]
            out = subprocess.run(["addr2line", "-f", "-C", "-e", binary] + rel, capture_output=True,
                                 text=True).stdout.splitlines()
            for i, (a, _) in enumerate(items):
                fn = out[2 * i] if 2 * i < len(out) else "??"
                name[a] = re.sub(r"\(.*", "", fn)[:100]
        else:
            addrs, names = dynsyms(f)
            short = f.split("/")[-1]
            for a, r in items:
                i = bisect.bisect_right(addrs, r) - 1
                name[a] = f"{names[i]} [{short}]" if i >= 0 else f"?? [{short}]"
    per_thread = collections.Counter(tids)
    print("-- CPU by thread: " + ", ".join(f"{t}={c}" for t, c in per_thread.most_common(8)))
    for t, c in per_thread.most_common(3):
        self_t, incl = collections.Counter(), collections.Counter()
        for s, tt in zip(stacks, tids):
            if tt != t:
                continue
            names = [name[a] for a in s]
            inner = next((n for n in names if not n.startswith(SIGNAL_FRAMES)), names[-1])
            self_t[inner] += 1
            for fn in set(names):
                incl[fn] += 1
        print(f"\n== thread {t}: {c} samples\n-- self")
        for fn, k in self_t.most_common(top):
            print(f"   {100.0 * k / c:6.2f}%  {fn}")
        print("-- inclusive")
        for fn, k in incl.most_common(top):
            print(f"   {100.0 * k / c:6.2f}%  {fn}")


if __name__ == "__main__":
    main()
