"""Build the UNMODIFIED reference (bytedance/ps-lite) for the baseline arm of bench.py.

`pip install /root/reference` is impossible (no setup.py / pyproject.toml — it is a C++
library built by a Makefile that wget's zeromq), so this script does what SURVEY §7.6
describes: copy the sources to baseline/_ref/src (git-ignored; /root/reference is
read-only), compile src/{customer,postoffice,van}.cc + tests/test_benchmark.cc with the
declaration shim in baseline/zmq_shim and link against the libzmq.so.5 bundled in pyzmq.
Only the ZMQ van can be built here (no ibverbs / UCX / libfabric in the image), so the
reference arm runs CPU buffers over TCP loopback or ipc:// (DMLC_LOCAL=1).
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "_ref")


def find_libzmq() -> str | None:
    import importlib.util

    spec = importlib.util.find_spec("zmq")
    if spec is None or not spec.origin:
        return None
    site = os.path.dirname(os.path.dirname(spec.origin))
    hits = sorted(glob.glob(os.path.join(site, "pyzmq.libs", "libzmq*.so*")))
    return hits[0] if hits else None


def build(force: bool = False) -> dict:
    out_bin = os.path.join(DST, "bin", "test_benchmark")
    if os.path.exists(out_bin) and not force:
        return {"ok": True, "bin": out_bin, "cached": True}
    src = os.path.join(DST, "src")
    if not os.path.isdir(src):
        if not os.path.isdir(REF_SRC):
            return {"ok": False, "why": f"reference sources not found at {REF_SRC}"}
        os.makedirs(DST, exist_ok=True)
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git"))
    libzmq = find_libzmq()
    if libzmq is None:
        return {"ok": False, "why": "no libzmq shared object found (pyzmq missing)"}
    os.makedirs(os.path.join(DST, "bin"), exist_ok=True)
    os.makedirs(os.path.join(DST, "lib"), exist_ok=True)
    # keep the original file name: it is the SONAME the binary will ask for
    local_zmq = os.path.join(DST, "lib", os.path.basename(libzmq))
    if not os.path.exists(local_zmq):
        shutil.copy(libzmq, local_zmq)
    # libzmq's own dependencies (libsodium, ...) ship next to it in pyzmq.libs
    deps = []
    for dep in glob.glob(os.path.join(os.path.dirname(libzmq), "*.so*")):
        if os.path.basename(dep).startswith("libzmq"):
            continue
        tgt = os.path.join(DST, "lib", os.path.basename(dep))
        if not os.path.exists(tgt):
            shutil.copy(dep, tgt)
        deps.append(tgt)
    srcs = [os.path.join(src, "src", f) for f in ("customer.cc", "postoffice.cc", "van.cc")]
    srcs.append(os.path.join(src, "tests", "test_benchmark.cc"))
    cmd = ["g++", "-std=c++14", "-O3", "-fopenmp", "-w", "-DDMLC_USE_ZMQ=1",
           f"-I{os.path.join(src, 'include')}", f"-I{os.path.join(src, 'src')}",
           f"-I{os.path.join(ROOT, 'zmq_shim')}", *srcs, local_zmq, *deps,
           f"-Wl,-rpath,{os.path.join(DST, 'lib')}", "-lpthread", "-lrt", "-o", out_bin]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return {"ok": False, "why": "compile failed: " + r.stderr.strip().splitlines()[-1][:200],
                "log": r.stderr}
    return {"ok": True, "bin": out_bin, "cached": False}


if __name__ == "__main__":
    res = build(force="--force" in sys.argv)
    print({k: v for k, v in res.items() if k != "log"})
    if not res["ok"] and "log" in res:
        print(res["log"][-3000:])
    sys.exit(0 if res["ok"] else 1)
