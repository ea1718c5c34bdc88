#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -s -k "native_symmetric or ipc_benchmark or stress_gpu" 2>&1 | tail -n 40
