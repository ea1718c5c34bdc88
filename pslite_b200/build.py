"""In-tree build of the native runtime (libpslite.a + apps) and the torch extension (_C.so).

`python -m pslite_b200.build` (or `build()` from `__graft_entry__`) cross-compiles every
CUDA source for sm_100a with `-gencode arch=compute_100a,code=sm_100a -lineinfo` (see the
Makefile) and links the PyTorch binding against the static runtime. Nothing is JIT-cached
under ~/.cache: the artefacts live in the tree so they travel to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT = os.path.join(ROOT, "pslite_b200", "_C.so")
BINDING = os.path.join(ROOT, "pslite_b200", "csrc", "bindings.cc")


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_native(jobs: int | None = None, verbose: bool = False) -> None:
    jobs = jobs or max(2, (os.cpu_count() or 4))
    cmd = ["make", "-j", str(jobs), "all"]
    subprocess.run(cmd, cwd=ROOT, check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    import torch
    from torch.utils.cpp_extension import include_paths, library_paths

    lib = os.path.join(ROOT, "build", "libpslite.a")
    if not force and _newer(EXT, BINDING, lib):
        return EXT
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    incs = include_paths("cuda") if hasattr(torch.utils.cpp_extension, "include_paths") else []
    try:
        incs = include_paths(device_type="cuda")
    except TypeError:
        incs = include_paths(cuda=True)
    incs += [sysconfig.get_paths()["include"], os.path.join(ROOT, "include"),
             os.path.join(ROOT, "src"), os.path.join(cuda_home, "include")]
    try:
        libs = library_paths(device_type="cuda")
    except TypeError:
        libs = library_paths(cuda=True)
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w",
           "-DPS_USE_CUDA=1", "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
    cmd += [f"-I{p}" for p in incs]
    cmd += [BINDING, lib]
    cmd += [f"-L{p}" for p in libs + [torch_lib, os.path.join(cuda_home, "lib64")]]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            "-lcudart", "-lrt", "-ldl", f"-Wl,-rpath,{torch_lib}", "-o", EXT]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, cwd=ROOT, check=True)
    return EXT


def build(force: bool = False, verbose: bool = False) -> None:
    build_native(verbose=verbose)
    build_extension(force=force, verbose=verbose)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", EXT)
