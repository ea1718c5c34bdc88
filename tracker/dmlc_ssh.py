#!/usr/bin/env python3
"""Launch a PS job over ssh on the hosts listed in a hostfile (one host[:port] per line).

    dmlc_ssh.py -n W -s S -H hosts command ...

Workers and servers are dealt round-robin over the hosts; the working directory can be
rsync'ed first with --sync-dir. Parity: reference tracker/dmlc_ssh.py:16-91.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from threading import Thread

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tracker  # noqa: E402

FORWARDED = ("OMP_NUM_THREADS", "KMP_AFFINITY", "LD_LIBRARY_PATH", "PS_VAN_TYPE", "PS_VERBOSE",
             "DMLC_INTERFACE", "DMLC_ENABLE_RDMA")


def read_hosts(path: str) -> list[tuple[str, str]]:
    hosts = []
    with open(path) as f:
        for line in f:
            h = line.strip()
            if not h or h.startswith("#"):
                continue
            host, _, port = h.partition(":")
            hosts.append((host, port or "22"))
    assert hosts, "empty hostfile"
    return hosts


def export_block(envs: dict) -> str:
    return "".join(f"export {k}={v}; " for k, v in envs.items())


def main():
    ap = argparse.ArgumentParser(description="run a pslite_b200 job over ssh")
    ap.add_argument("-n", "--num-workers", required=True, type=int)
    ap.add_argument("-s", "--num-servers", default=0, type=int)
    ap.add_argument("-H", "--hostfile", required=True)
    ap.add_argument("--sync-dir", default=None, help="rsync the current directory there first")
    ap.add_argument("--host-ip", default="auto")
    ap.add_argument("--dry-run", action="store_true", help="print the ssh commands only")
    ap.add_argument("command", nargs="+")
    args, unknown = ap.parse_known_args()
    cmd = " ".join(args.command + unknown)
    hosts = read_hosts(args.hostfile)
    cwd = os.getcwd()
    if args.sync_dir:
        for host, port in hosts:
            sync = ["rsync", "-az", "--rsh", f"ssh -o StrictHostKeyChecking=no -p {port}",
                    cwd + "/", f"{host}:{args.sync_dir}"]
            print(" ".join(sync)) if args.dry_run else subprocess.check_call(sync)
        cwd = args.sync_dir

    def submit(nworker, nserver, pass_envs):
        envs = dict(pass_envs)
        for k in FORWARDED:
            if k in os.environ:
                envs[k] = os.environ[k]
        threads = []
        for i in range(nworker + nserver):
            envs["DMLC_ROLE"] = "server" if i < nserver else "worker"
            host, port = hosts[i % len(hosts)]
            remote = f"{export_block(envs)} cd {cwd}; {cmd}"
            ssh = f"ssh -o StrictHostKeyChecking=no {host} -p {port} '{remote}'"
            if args.dry_run:
                print(ssh)
                continue
            t = Thread(target=lambda c=ssh: subprocess.check_call(c, shell=True), daemon=True)
            t.start()
            threads.append(t)
        for t in threads:
            t.join()

    if args.dry_run:
        submit(args.num_workers, args.num_servers,
               {"DMLC_NUM_WORKER": args.num_workers, "DMLC_NUM_SERVER": args.num_servers,
                "DMLC_PS_ROOT_URI": "<tracker-ip>", "DMLC_PS_ROOT_PORT": "<port>"})
        return
    tracker.config_logger()
    tracker.submit(args.num_workers, args.num_servers, fun_submit=submit, host_ip=args.host_ip,
                   pscmd=cmd)


if __name__ == "__main__":
    main()
