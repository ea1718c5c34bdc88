#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -n 3
