#!/bin/bash
# Push/pull preset of the reference's two-node script (test.sh:2,21-27: 4,096,000 B, mode 1),
# here on one box: `local` = TCP van, `nvl` = values in HBM over the NVLink van (1 GPU per process).
cd "$(dirname "$0")/.."
mode=${1:-local}
export BENCHMARK_NTHREAD=${BENCHMARK_NTHREAD:-1} TOTAL_DURATION=${TOTAL_DURATION:-100} LOG_DURATION=${LOG_DURATION:-20}
if [ "$mode" = "nvl" ]; then
  export PS_VAN_TYPE=nvl TEST_NUM_GPU_WORKER=1 TEST_NUM_GPU_SERVER=1 WORKER_GPU_BASE=0 SERVER_GPU_BASE=${NUM_WORKERS:-1}
fi
exec scripts/local.sh ${NUM_SERVERS:-1} ${NUM_WORKERS:-1} build/test_benchmark 4096000 100000 1
