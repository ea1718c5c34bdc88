#!/usr/bin/env python3
"""Launch a PS job on this machine: 1 scheduler + S servers + W workers.

    dmlc_local.py -n W -s S [--van nvl|zmq|shm] [--gpus-per-role] command ...

Each process inherits the DMLC_* rendezvous variables; a process that exits with code 254
is restarted (DMLC_NUM_ATTEMPT counts the attempts). With --gpus, worker i gets
PS_CUDA_DEVICE=i and server j gets PS_CUDA_DEVICE=W+j (one GPU per process).
Parity: reference tracker/dmlc_local.py:15-25 (keepalive), :27-77 (LocalLauncher).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from threading import Thread

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tracker  # noqa: E402


class LocalLauncher:
    def __init__(self, args, unknown):
        self.args = args
        self.cmd = " ".join(args.command + unknown)
        self.failures: list[int] = []

    def _keepalive(self, role: str, env: dict) -> None:
        attempt = 0
        while True:
            e = dict(env)
            e["DMLC_NUM_ATTEMPT"] = str(attempt)
            rc = subprocess.call(self.cmd, shell=True, env=e)
            if rc == 254:  # "restart me"
                attempt += 1
                continue
            if rc != 0:
                self.failures.append(rc)
            return

    def submit(self, nworker: int, nserver: int, envs: dict) -> None:
        threads = []
        for i in range(nworker + nserver):
            role = "worker" if i < nworker else "server"
            env = os.environ.copy()
            env.update({k: str(v) for k, v in envs.items()})
            env["DMLC_ROLE"] = role
            if self.args.van:
                env["PS_VAN_TYPE"] = self.args.van
            if self.args.gpus:
                env["PS_CUDA_DEVICE"] = str(i)
            t = Thread(target=self._keepalive, args=(role, env), daemon=True)
            t.start()
            threads.append(t)
        for t in threads:
            while t.is_alive():
                t.join(100)

    def run(self) -> int:
        tracker.config_logger(self.args.log_level)
        tracker.submit(self.args.num_workers, self.args.num_servers, fun_submit=self.submit,
                       host_ip=self.args.host_ip, pscmd=self.cmd)
        return 1 if self.failures else 0


def main():
    ap = argparse.ArgumentParser(description="run a pslite_b200 job on the local machine")
    ap.add_argument("-n", "--num-workers", required=True, type=int)
    ap.add_argument("-s", "--num-servers", type=int, default=0)
    ap.add_argument("--host-ip", default="127.0.0.1")
    ap.add_argument("--van", default=None, help="PS_VAN_TYPE for every process")
    ap.add_argument("--gpus", action="store_true", help="pin process i to GPU i")
    ap.add_argument("--log-level", default="INFO", choices=["INFO", "DEBUG"])
    ap.add_argument("command", nargs="+")
    args, unknown = ap.parse_known_args()
    sys.exit(LocalLauncher(args, unknown).run())


if __name__ == "__main__":
    main()
