"""Turn a torchrun job (one process per GPU) into a parameter-server cluster.

torchrun gives RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT; ps-lite wants a
scheduler plus DMLC_* roles (reference docs/env.md:5-15, tracker/dmlc_local.py:59-70).
Rank 0 spawns the scheduler as a child process; every rank then starts its role(s):

    topology "joint": every rank is worker r AND server r (co-located, the IPC-benchmark
                      shape of the reference, tests/test_ipc_benchmark.cc) — all N GPUs compute.
    topology "split": ranks [0, N/2) are workers, ranks [N/2, N) are servers
                      (the 4w+4s shape named in BASELINE.json).
"""
from __future__ import annotations

import os
import subprocess
import sys
from dataclasses import dataclass

from ..utils.env import free_port


@dataclass
class PSContext:
    rank: int
    world: int
    local_rank: int
    topology: str
    num_workers: int
    num_servers: int
    is_worker: bool
    is_server: bool
    worker_rank: int
    server_rank: int
    role: str
    van: str
    scheduler: subprocess.Popen | None = None

    def shutdown(self):
        """Normal end of the job: barrier with every node, stop the vans, reap the scheduler."""
        from .. import native

        native().finalize(0, self.role, True)
        if self.scheduler is not None:
            try:
                self.scheduler.wait(timeout=60)
            except subprocess.TimeoutExpired:
                self.scheduler.kill()
            self.scheduler = None

    def abort(self):
        """Error path: do not wait for peers; make sure no scheduler child outlives us (an
        orphan would also keep inherited stdout/stderr pipes open for whoever waits on them)."""
        if self.scheduler is not None:
            self.scheduler.kill()
            self.scheduler = None


def _share_port(rank: int, world: int) -> int:
    """Pick the scheduler port on rank 0 and tell everyone (torch.distributed if present)."""
    if world == 1:
        return free_port()
    import torch.distributed as dist

    assert dist.is_initialized(), "init torch.distributed before init_ps() when WORLD_SIZE > 1"
    box = [free_port() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return int(box[0])


def bind_to_gpu_numa_node(device_index: int) -> list[int] | None:
    """Restrict this process (and the threads it starts later: van receive threads, the outbox, torch's
    helpers) to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers are first-touched
    on that node and host<->device copies do not cross the socket interconnect. Eight ranks started by
    torchrun otherwise land on arbitrary CPUs. Returns the CPU list, or None if nothing was changed
    (no sysfs entry, a node with fewer than 8 CPUs, PS_NUMA_BIND=0)."""
    if os.environ.get("PS_NUMA_BIND", "1") == "0":
        return None
    try:
        import torch

        if not torch.cuda.is_available():
            return None
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus: list[int] = []
        for part in spec.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if len(cpus) < 8 or len(cpus) >= len(allowed):
            # no NUMA information worth acting on (or already bound); a rank keeps three threads busy (the
            # issuing thread and one receive thread per van), so a node with a handful of CPUs is no home
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # a missing attribute / sysfs file must never stop a job
        return None


def init_ps(topology: str = "joint", van: str | None = None, extra_env: dict | None = None) -> PSContext:
    """Start the PS runtime for this torchrun rank and return its context."""
    from .. import native

    C = native()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    host = os.environ.get("MASTER_ADDR", "127.0.0.1")
    if van is None:
        van = "nvl" if C.cuda_usable() else "shm"
    if topology == "split" and world < 2:
        topology = "joint"
    if topology == "joint":
        nw = ns = world
        is_worker = is_server = True
        wrank = srank = rank
        role = "joint"
    elif topology == "split":
        assert world % 2 == 0, "split topology needs an even number of ranks"
        nw = ns = world // 2
        is_worker = rank < nw
        is_server = not is_worker
        wrank = rank if is_worker else -1
        srank = rank - nw if is_server else -1
        role = "worker" if is_worker else "server"
    else:
        raise ValueError(f"unknown topology {topology}")

    port = _share_port(rank, world)
    env = {
        "DMLC_NUM_WORKER": nw, "DMLC_NUM_SERVER": ns,
        "DMLC_PS_ROOT_URI": host if host not in ("localhost",) else "127.0.0.1",
        "DMLC_PS_ROOT_PORT": port,
        "PS_VAN_TYPE": van, "PS_CUDA_DEVICE": local_rank, "DMLC_ROLE": role,
    }
    # one node: everything on loopback (the container's hostname may not resolve). Several nodes
    # (torchrun --nnodes > 1): every process advertises its own address — DMLC_NODE_HOST if the
    # user set it, else the van picks the first non-loopback interface (or DMLC_INTERFACE).
    single_node = int(os.environ.get("LOCAL_WORLD_SIZE", str(world))) >= world
    if single_node:
        env["DMLC_NODE_HOST"] = "127.0.0.1"
        if os.environ.get("PS_TEST_HOST_PER_RANK"):
            # tests: every rank announces another loopback address, so the vans take the other ranks for
            # other hosts (frames instead of one-sided writes, device values staged through the host)
            env["DMLC_NODE_HOST"] = f"127.0.0.{rank + 1}"
    elif os.environ.get("DMLC_NODE_HOST"):
        env["DMLC_NODE_HOST"] = os.environ["DMLC_NODE_HOST"]
    if van == "nvl":
        # one process per GPU: raw copies without a producer event go through the copy engine (an
        # on-demand persistent kernel fed from a host-mapped ring) instead of one launch each
        env["PS_COPY_ENGINE"] = os.environ.get("PS_COPY_ENGINE", "1")
    if extra_env:
        env.update(extra_env)
    sched = None
    if rank == 0:
        child_env = dict(os.environ)
        child_env.update({k: str(v) for k, v in env.items()})
        child_env["DMLC_ROLE"] = "scheduler"
        child_env["PS_VAN_TYPE"] = "zmq"  # the scheduler moves no payload
        sched = subprocess.Popen([sys.executable, "-m", "pslite_b200.parallel.scheduler"],
                                 env=child_env)
        import atexit

        atexit.register(lambda p=sched: p.poll() is None and p.kill())
    for k, v in env.items():
        C.set_env(k, str(v))
    if van == "nvl":
        bind_to_gpu_numa_node(local_rank)  # before the van starts its threads and anybody pins host memory
    preferred = wrank if is_worker else srank
    C.start_ps(0, role, preferred, True)
    return PSContext(rank, world, local_rank, topology, nw, ns, is_worker, is_server, wrank, srank,
                     role, van, sched)
