#!/usr/bin/env python3
"""Launch a PS job with mpirun (Open MPI `-x` or MPICH `-env` style env forwarding).

    dmlc_mpi.py -n W -s S [-H hostfile] command ...
Parity: reference tracker/dmlc_mpi.py:33-91.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from threading import Thread

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tracker  # noqa: E402


def mpi_env_flags(envs: dict) -> str:
    try:
        out = subprocess.run(["mpirun", "--version"], capture_output=True, text=True).stdout
    except OSError:
        out = ""
    if "Open MPI" in out or "OpenRTE" in out:
        return " ".join(f"-x {k}={v}" for k, v in envs.items())
    return " ".join(f"-env {k} {v}" for k, v in envs.items())


def main():
    ap = argparse.ArgumentParser(description="run a pslite_b200 job with MPI")
    ap.add_argument("-n", "--num-workers", required=True, type=int)
    ap.add_argument("-s", "--num-servers", default=0, type=int)
    ap.add_argument("-H", "--hostfile", default=None)
    ap.add_argument("--host-ip", default="auto")
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("command", nargs="+")
    args, unknown = ap.parse_known_args()
    cmd = " ".join(args.command + unknown)

    def submit(nworker, nserver, pass_envs):
        threads = []
        for role, n in (("server", nserver), ("worker", nworker)):
            if n == 0:
                continue
            envs = dict(pass_envs, DMLC_ROLE=role)
            host = f"--hostfile {args.hostfile}" if args.hostfile else ""
            line = f"mpirun -n {n} {host} {mpi_env_flags(envs)} {cmd}"
            if args.dry_run:
                print(line)
                continue
            t = Thread(target=lambda c=line: subprocess.check_call(c, shell=True), daemon=True)
            t.start()
            threads.append(t)
        for t in threads:
            t.join()

    if args.dry_run:
        submit(args.num_workers, args.num_servers,
               {"DMLC_NUM_WORKER": args.num_workers, "DMLC_NUM_SERVER": args.num_servers})
        return
    tracker.config_logger()
    tracker.submit(args.num_workers, args.num_servers, fun_submit=submit, host_ip=args.host_ip,
                   pscmd=cmd)


if __name__ == "__main__":
    main()
