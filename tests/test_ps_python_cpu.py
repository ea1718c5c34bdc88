"""Python-level push/pull through the native runtime on CPU (tcp and shm vans)."""
import os
import subprocess
import sys

import pytest

from pslite_b200.utils.env import free_port

HERE = os.path.dirname(os.path.abspath(__file__))
NODE = os.path.join(HERE, "helpers", "ps_node.py")


def run_cluster(van, nw, ns, n_elems=1000, timeout=120):
    port = str(free_port())
    procs = []
    env = dict(os.environ)
    env["PSLITE_NO_AUTOBUILD"] = "1"
    for role, count in (("scheduler", 1), ("server", ns), ("worker", nw)):
        for _ in range(count):
            procs.append((role, subprocess.Popen(
                [sys.executable, NODE, role, van, str(nw), str(ns), port, str(n_elems)],
                env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    outs = []
    rc = 0
    for role, p in procs:
        try:
            o, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for _, q in procs:
                q.kill()
            o, _ = p.communicate()
            rc = rc or 124
        outs.append((role, p.returncode, o))
        rc = rc or p.returncode
    return rc, outs


@pytest.mark.parametrize("van,nw,ns", [("zmq", 1, 1), ("zmq", 2, 2), ("shm", 1, 1), ("shm", 2, 2)])
def test_python_push_pull(native, van, nw, ns):
    rc, outs = run_cluster(van, nw, ns)
    assert rc == 0, "\n".join(f"[{r} rc={c}]\n{o[-1500:]}" for r, c, o in outs)
    passes = sum(o.count("PASS") for r, c, o in outs if r == "worker")
    assert passes == nw


def test_symmetric_memory_shm(native):
    """Van::AllocSymmetric on the shm van: 2 workers + 2 servers (4 processes) allocate one symmetric
    buffer through FdExchange (descriptors over unix sockets) and read each other's blocks"""
    port = str(free_port())
    env = dict(os.environ)
    env["PSLITE_NO_AUTOBUILD"] = "1"
    helper = os.path.join(HERE, "helpers", "symm_node.py")
    procs = [(role, subprocess.Popen([sys.executable, helper, role, "shm", "2", "2", port], env=env,
                                     stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
             for role in ("scheduler", "server", "server", "worker", "worker")]
    outs = []
    for role, p in procs:
        try:
            o, _ = p.communicate(timeout=120)
        except subprocess.TimeoutExpired:
            for _, q in procs:
                q.kill()
            o, _ = p.communicate()
        outs.append((role, p.returncode, o))
    assert all(c == 0 for _, c, _ in outs), "\n".join(f"[{r} rc={c}]\n{o[-1500:]}" for r, c, o in outs)
    assert sum(o.count("PASS") for _, _, o in outs) == 4


def test_python_large_message_shm(native):
    rc, outs = run_cluster("shm", 1, 1, n_elems=2_000_000)
    assert rc == 0, "\n".join(f"[{r} rc={c}]\n{o[-1500:]}" for r, c, o in outs)


def test_trainer_logic_on_cpu(native):
    """PSWorkerOptimizer (hooks, chunked keys, init, sync rounds) against a Python SGD server."""
    node = os.path.join(HERE, "helpers", "train_cpu_node.py")
    port = str(free_port())
    env = dict(os.environ)
    env["PSLITE_NO_AUTOBUILD"] = "1"
    nw, steps = 2, 5
    procs = []
    for role, count in (("scheduler", 1), ("server", 1), ("worker", nw)):
        for _ in range(count):
            procs.append((role, subprocess.Popen([sys.executable, node, role, str(nw), port, str(steps)],
                                                 env=env, stdout=subprocess.PIPE,
                                                 stderr=subprocess.STDOUT, text=True)))
    outs = []
    for role, p in procs:
        try:
            o, _ = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            for _, q in procs:
                q.kill()
            o, _ = p.communicate()
        outs.append((role, p.returncode, o))
    assert all(c == 0 for _, c, _ in outs), "\n".join(f"[{r} rc={c}]\n{o[-2000:]}" for r, c, o in outs)
    sums = []
    for role, _, o in outs:
        for line in o.splitlines():
            if line.startswith("CHECKSUM"):
                f = line.split()
                sums.append((float(f[2]), float(f[4])))
    assert len(sums) == nw
    assert abs(sums[0][0] - sums[1][0]) < 1e-9          # all workers hold the same parameters
    assert abs(sums[0][0] - sums[0][1]) < 5e-2          # and they match the local simulation


@pytest.mark.parametrize("nproc,topo,length,async_copies", [
    (8, "split", 4096, "1"), (4, "split", 1 << 20, "1"), (4, "joint", 65536, "1"), (4, "split", 4096, "0")])
def test_bench_loop_many_peers(native, nproc, topo, length, async_copies, coalesce="0", fused="0", staged="0"):
    """bench.py's push_pull_batch loop under torchrun with several workers AND servers, over the
    one-sided van with *asynchronous* copies (PS_SHM_ASYNC: copies complete later, as kernels
    on a CUDA stream do). Descriptors for different peers then share completion batches — the
    case a per-peer batch of one once mishandled (pull replies arrived without their MemRef)."""
    helper = os.path.join(HERE, "helpers", "pushpull_multi.py")
    env = dict(os.environ)
    env.update({"PSLITE_NO_AUTOBUILD": "1", "PS_SHM_ASYNC": async_copies, "OMP_NUM_THREADS": "1",
                "PS_COALESCE_LAUNCHES": coalesce, "PSLITE_TEST_PUSHPULL": fused, "PSLITE_TEST_STAGED": staged})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), helper, "shm", str(length), "10",
           "10", topo]
    for attempt in range(2 if nproc == 8 and coalesce == "0" and fused == "0" and staged == "0" else 1):  # the hang was probabilistic
        p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=150)
        assert p.returncode == 0 and "PASS" in p.stdout, (p.stdout + p.stderr)[-3000:]


def test_bench_loop_with_launch_coalescing(native):
    """the same loop with every push_pull_batch call and every handler batch corked"""
    test_bench_loop_many_peers(native, 8, "split", 65536, "1", coalesce="1")


def test_bench_loop_with_fused_push_pull(native):
    """one KVWorker::ZPushPull per key (request carries the push, the single reply the pulled values)"""
    test_bench_loop_many_peers(native, 8, "split", 65536, "1", fused="1")


def test_bench_loop_with_staged_rounds(native):
    """KVWorker.staged_push_pull: host -> tensor -> push -> pull -> host for every key in one call"""
    test_bench_loop_many_peers(native, 4, "split", 65536, "1", staged="1")
