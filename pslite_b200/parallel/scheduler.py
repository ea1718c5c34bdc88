"""Scheduler process: registers nodes, assigns ranks, brokers barriers, then exits.

Equivalent of running the reference binary with DMLC_ROLE=scheduler
(tests/test_benchmark.cc:524,551: the scheduler only does StartPS / Finalize).
"""
from __future__ import annotations


def _die_with_parent() -> None:
    """A scheduler whose launcher crashed would wait for its nodes forever (and keep the port and
    the launcher's stdout pipe): ask the kernel for SIGTERM when the parent goes away."""
    import ctypes
    import os
    import signal

    parent = os.getppid()
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        libc.prctl(1, signal.SIGTERM, 0, 0, 0)  # PR_SET_PDEATHSIG
    except (OSError, AttributeError):
        return
    if os.getppid() != parent:  # the parent died before prctl took effect
        os.kill(os.getpid(), signal.SIGTERM)


def main() -> None:
    import os

    if os.environ.get("PS_SCHEDULER_ORPHAN_OK", "0") != "1":
        _die_with_parent()
    from pslite_b200 import native

    C = native()
    C.start_ps(0, "scheduler", -1, True)
    C.finalize(0, "scheduler", True)


if __name__ == "__main__":
    main()
