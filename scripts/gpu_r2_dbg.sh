#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
for i in 1 2 3 4; do
  env PS_VAN_TYPE=nvl JOINT=1 WORKER_GPU_BASE=0 BENCHMARK_NTHREAD=2 PS_DROP_MSG=5 PS_RESEND=1 PS_RESEND_TIMEOUT=200 \
    timeout 150 scripts/local.sh 2 2 build/test_benchmark_stress 4096000 4 > gpurun_out/stress_$i.log 2>&1
  echo "run $i rc=$?"; grep -a "PASSED\|FAILED\|Check failed\| F [a-z_]*\.\|corrupt\|what()" gpurun_out/stress_$i.log | cut -c1-400 | head -8
done
echo "== several GPUs per process test + others"
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q -s -k "several_gpus" 2>&1 | tail -n 25 | cut -c1-300
