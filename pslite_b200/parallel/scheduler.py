"""Scheduler process: registers nodes, assigns ranks, brokers barriers, then exits.

Equivalent of running the reference binary with DMLC_ROLE=scheduler
(tests/test_benchmark.cc:524,551: the scheduler only does StartPS / Finalize).
"""
from __future__ import annotations


def main() -> None:
    from pslite_b200 import native

    C = native()
    C.start_ps(0, "scheduler", -1, True)
    C.finalize(0, "scheduler", True)


if __name__ == "__main__":
    main()
