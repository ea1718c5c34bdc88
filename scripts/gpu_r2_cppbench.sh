#!/bin/bash
# apps/test_benchmark on the nvl van, 1 worker GPU + 1 server GPU, with the issue / wait split per
# key; engine off / on; 1 KB ... 16 MB messages.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PS_VAN_TYPE=nvl TEST_NUM_GPU_WORKER=1 TEST_NUM_GPU_SERVER=1 NUM_KEY_PER_SERVER=40
for eng in 0 1; do
  for sz in 1024 65536 1048576 4096000 16777216; do
    rounds=2000; [ $sz -gt 100000 ] && rounds=400
    echo "== split 2 GPUs, engine=$eng, $sz B"
    WORKER_GPU_BASE=0 SERVER_GPU_BASE=1 PS_COPY_ENGINE=$eng TOTAL_DURATION=$rounds LOG_DURATION=$((rounds/4)) PS_VERBOSE=1 ${EXTRA_ENV} \
      timeout 120 scripts/local.sh 1 1 build/test_benchmark $sz 100 1 > gpurun_out/cpp_tmp.log 2>&1
    grep -i "goodput" gpurun_out/cpp_tmp.log | tail -n 2 | cut -c30-260
    grep -i "gated" gpurun_out/cpp_tmp.log | grep -v " 0 one-sided" | cut -c30-200 | head -n 3
    grep -i "error\|fatal\|Check failed" gpurun_out/cpp_tmp.log | head -n 3
  done
done
