"""Job trackers for pslite_b200 (Python 3).

* PSTracker   — starts the scheduler process of a parameter-server job and hands the
                DMLC_* rendezvous environment to the worker / server launchers.
* RabitTracker — rendezvous broker for rabit-style allreduce workers: accepts worker
                connections, assigns ranks and tells each worker its tree parent/children
                and ring neighbours. Not used by the PS itself; kept because the
                reference's launcher serves both kinds of jobs.
* submit()    — glue used by the dmlc_local / dmlc_ssh / dmlc_mpi front ends.

Parity: reference tracker/tracker.py:127-316 (RabitTracker), :318-365 (PSTracker),
:367-420 (submit). Re-written for Python 3 (the reference is Python 2: xrange, str sockets).
"""
from __future__ import annotations

import logging
import os
import socket
import struct
import subprocess
import sys
import time
from threading import Thread

MAGIC = 0xFF99


class _Wire:
    """length-prefixed ints / strings over a TCP socket (rabit tracker protocol)"""

    def __init__(self, sock: socket.socket):
        self.sock = sock

    def recvall(self, n: int) -> bytes:
        chunks, got = [], 0
        while got < n:
            c = self.sock.recv(min(n - got, 1024))
            if not c:
                raise ConnectionError("peer closed")
            chunks.append(c)
            got += len(c)
        return b"".join(chunks)

    def recvint(self) -> int:
        return struct.unpack("@i", self.recvall(4))[0]

    def sendint(self, v: int) -> None:
        self.sock.sendall(struct.pack("@i", v))

    def sendstr(self, s: str) -> None:
        b = s.encode()
        self.sendint(len(b))
        self.sock.sendall(b)

    def recvstr(self) -> str:
        return self.recvall(self.recvint()).decode()


def _resolve_ip(host: str) -> str:
    return socket.gethostbyname(host) if host else "127.0.0.1"


def get_host_ip(host_ip: str | None = None) -> str:
    """Best-effort address other machines can reach this host on."""
    if host_ip in (None, "auto", "ip"):
        try:
            ip = socket.gethostbyname(socket.getfqdn())
        except OSError:
            ip = "127.0.0.1"
        if ip.startswith("127."):
            s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
            try:
                s.connect(("10.255.255.255", 1))  # no packet is sent
                ip = s.getsockname()[0]
            except OSError:
                ip = "127.0.0.1"
            finally:
                s.close()
        return ip
    if host_ip == "dns":
        return socket.getfqdn()
    return host_ip


def bind_free_port(host: str, port: int, port_end: int) -> tuple[socket.socket, int]:
    sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    for p in range(port, port_end):
        try:
            sock.bind((host, p))
            return sock, p
        except OSError:
            continue
    raise OSError(f"no free port in [{port}, {port_end})")


# ---------------------------------------------------------------------------------------
# topology helpers (binary tree + ring that shares as many links with the tree as possible)
# ---------------------------------------------------------------------------------------
def tree_neighbours(rank: int, n: int) -> list[int]:
    r = rank + 1
    out = []
    if r > 1:
        out.append(r // 2 - 1)
    if r * 2 - 1 < n:
        out.append(r * 2 - 1)
    if r * 2 < n:
        out.append(r * 2)
    return out


def build_tree(n: int) -> tuple[dict[int, list[int]], dict[int, int]]:
    tree = {r: tree_neighbours(r, n) for r in range(n)}
    parent = {r: (r + 1) // 2 - 1 for r in range(n)}
    return tree, parent


def dfs_ring_order(tree, parent, root: int = 0) -> list[int]:
    """visit order of a DFS over the tree: consecutive entries are (mostly) tree neighbours"""
    order, stack = [], [root]
    while stack:
        r = stack.pop()
        order.append(r)
        kids = [c for c in tree[r] if c != parent[r]]
        stack.extend(reversed(kids))
    return order


def build_ring(tree, parent) -> dict[int, tuple[int, int]]:
    order = dfs_ring_order(tree, parent)
    n = len(order)
    return {order[i]: (order[(i - 1) % n], order[(i + 1) % n]) for i in range(n)}


def link_map(n: int):
    """(tree, parent, ring) relabelled so that ring order is 0,1,2,... (rabit's convention)"""
    tree, parent = build_tree(n)
    ring = build_ring(tree, parent)
    relabel = {0: 0}
    cur = 0
    for i in range(n - 1):
        cur = ring[cur][1]
        relabel[cur] = i + 1
    tree2 = {relabel[k]: [relabel[x] for x in v] for k, v in tree.items()}
    parent2 = {relabel[k]: (relabel[v] if k != 0 else -1) for k, v in parent.items()}
    ring2 = {relabel[k]: (relabel[v[0]], relabel[v[1]]) for k, v in ring.items()}
    return tree2, parent2, ring2


class _WorkerEntry:
    def __init__(self, sock: socket.socket, addr):
        self.wire = _Wire(sock)
        self.sock = sock
        self.host = _resolve_ip(addr[0])
        magic = self.wire.recvint()
        if magic != MAGIC:
            raise ConnectionError(f"invalid magic {magic:#x} from {self.host}")
        self.wire.sendint(MAGIC)
        self.rank = self.wire.recvint()
        self.world_size = self.wire.recvint()
        self.jobid = self.wire.recvstr()
        self.cmd = self.wire.recvstr()
        self.wait_accept = 0
        self.port = None

    def decide_rank(self, job_map: dict) -> int:
        if self.rank >= 0:
            return self.rank
        if self.jobid != "NULL" and self.jobid in job_map:
            return job_map[self.jobid]
        return -1

    def assign_rank(self, rank, wait_conn, tree, parent, ring) -> list[int]:
        self.rank = rank
        nnset = set(tree[rank])
        rprev, rnext = ring[rank]
        w = self.wire
        w.sendint(rank)
        w.sendint(parent[rank])
        w.sendint(len(tree))
        w.sendint(len(nnset))
        for r in nnset:
            w.sendint(r)
        w.sendint(rprev if rprev not in (-1, rank) else -1)
        w.sendint(rnext if rnext not in (-1, rank) else -1)
        if rprev not in (-1, rank):
            nnset.add(rprev)
        if rnext not in (-1, rank):
            nnset.add(rnext)
        while True:
            ngood = w.recvint()
            good = {w.recvint() for _ in range(ngood)}
            assert good.issubset(nnset)
            bad = nnset - good
            conset = [r for r in bad if r in wait_conn]
            w.sendint(len(conset))
            w.sendint(len(bad) - len(conset))
            for r in conset:
                w.sendstr(wait_conn[r].host)
                w.sendint(wait_conn[r].port)
                w.sendint(r)
            nerr = w.recvint()
            if nerr != 0:
                continue
            self.port = w.recvint()
            done = []
            for r in conset:
                wait_conn[r].wait_accept -= 1
                if wait_conn[r].wait_accept == 0:
                    done.append(r)
            for r in done:
                wait_conn.pop(r, None)
            self.wait_accept = len(bad) - len(conset)
            return done


class RabitTracker:
    """Rendezvous server for rabit workers (tree + ring allreduce topology)."""

    def __init__(self, host_ip: str, nworker: int, port: int = 9091, port_end: int = 9999):
        self.sock, self.port = bind_free_port(host_ip, port, port_end)
        self.sock.listen(256)
        self.host_ip = host_ip
        self.nworker = nworker
        self.thread = None
        self.start_time = None
        self.end_time = None
        logging.info("start listen on %s:%d", host_ip, self.port)

    def __del__(self):
        try:
            self.sock.close()
        except OSError:
            pass

    def worker_envs(self) -> dict:
        return {"DMLC_TRACKER_URI": self.host_ip, "DMLC_TRACKER_PORT": self.port}

    def _accept_workers(self, nworker: int) -> None:
        shutdown, wait_conn, job_map, pending = {}, {}, {}, []
        tree = parent = ring = None
        todo = None
        while len(shutdown) != nworker:
            fd, addr = self.sock.accept()
            s = _WorkerEntry(fd, addr)
            if s.cmd == "print":
                logging.info(s.wire.recvstr().strip())
                continue
            if s.cmd == "shutdown":
                assert s.rank >= 0 and s.rank not in shutdown and s.rank not in wait_conn
                shutdown[s.rank] = s
                continue
            assert s.cmd in ("start", "recover")
            if tree is None:
                assert s.cmd == "start"
                if s.world_size > 0:
                    nworker = s.world_size
                tree, parent, ring = link_map(nworker)
                todo = list(range(nworker))
            else:
                assert s.world_size in (-1, nworker)
            if s.cmd == "recover":
                assert s.rank >= 0
            rank = s.decide_rank(job_map)
            if rank == -1:
                assert todo
                pending.append(s)
                if len(pending) == len(todo):
                    pending.sort(key=lambda x: x.host)
                    for p in pending:
                        rank = todo.pop(0)
                        if p.jobid != "NULL":
                            job_map[p.jobid] = rank
                        p.assign_rank(rank, wait_conn, tree, parent, ring)
                        if p.wait_accept > 0:
                            wait_conn[rank] = p
                    pending = []
                if not todo:
                    logging.info("@tracker all of %d nodes getting started", nworker)
                    self.start_time = time.time()
            else:
                s.assign_rank(rank, wait_conn, tree, parent, ring)
                if s.wait_accept > 0:
                    wait_conn[rank] = s
        logging.info("@tracker all nodes finished the job")
        self.end_time = time.time()

    def start(self, nworker: int) -> None:
        self.thread = Thread(target=self._accept_workers, args=(nworker,), daemon=True)
        self.thread.start()

    def join(self) -> None:
        while self.thread is not None and self.thread.is_alive():
            self.thread.join(100)


class PSTracker:
    """Starts the scheduler of a PS job and publishes its address."""

    def __init__(self, host_ip: str, cmd: str | None, port: int = 9091, port_end: int = 9999,
                 envs: dict | None = None):
        self.cmd = cmd
        self.host_ip = host_ip
        self.thread = None
        if cmd is None:
            self.port = port
            return
        sock, self.port = bind_free_port("", port, port_end)
        sock.close()
        env = os.environ.copy()
        env["DMLC_ROLE"] = "scheduler"
        env["DMLC_PS_ROOT_URI"] = str(self.host_ip)
        env["DMLC_PS_ROOT_PORT"] = str(self.port)
        for k, v in (envs or {}).items():
            env[k] = str(v)
        self.thread = Thread(target=lambda: subprocess.check_call(self.cmd, env=env, shell=True),
                             daemon=True)
        self.thread.start()

    def join(self) -> None:
        if self.cmd is not None:
            while self.thread.is_alive():
                self.thread.join(100)

    def worker_envs(self) -> dict:
        if self.cmd is None:
            return {}
        return {"DMLC_PS_ROOT_URI": self.host_ip, "DMLC_PS_ROOT_PORT": self.port}


def submit(nworker: int, nserver: int, fun_submit, host_ip: str = "auto", pscmd: str | None = None):
    """Start the tracker(s), call `fun_submit(nworker, nserver, envs)` to launch the job,
    then wait for it. nserver == 0 means a rabit (allreduce) job."""
    envs = {"DMLC_NUM_WORKER": nworker, "DMLC_NUM_SERVER": nserver}
    host_ip = get_host_ip(host_ip)
    if nserver == 0:
        rabit = RabitTracker(host_ip=host_ip, nworker=nworker)
        envs.update(rabit.worker_envs())
        rabit.start(nworker)
        tracker = rabit
    else:
        tracker = PSTracker(host_ip=host_ip, cmd=pscmd, envs=envs)
        envs.update(tracker.worker_envs())
    fun_submit(nworker, nserver, envs)
    tracker.join()


def config_logger(level: str = "INFO") -> None:
    logging.basicConfig(format="%(asctime)s %(levelname)s %(message)s", level=getattr(logging, level))


if __name__ == "__main__":
    print("use dmlc_local.py / dmlc_ssh.py / dmlc_mpi.py", file=sys.stderr)
