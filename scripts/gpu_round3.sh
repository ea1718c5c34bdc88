#!/bin/bash
# 2-GPU validation: cross-process CUDA IPC + NVLink peer writes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for len in 1048576 67108864 268435456; do
  echo "== test_benchmark nvl 1w(GPU0)+1s(GPU1) len=$len"
  PS_VAN_TYPE=nvl TEST_NUM_GPU_WORKER=1 TEST_NUM_GPU_SERVER=1 WORKER_GPU_BASE=0 SERVER_GPU_BASE=1 NUM_KEY_PER_SERVER=8 \
    TOTAL_DURATION=40 LOG_DURATION=20 timeout 120 scripts/local.sh 1 1 build/test_benchmark $len 10 1 \
    > gpurun_out/tb_nvl2_$len.log 2>&1; echo "rc=$?"
  grep -E "goodput|Check failed|rror" gpurun_out/tb_nvl2_$len.log | head -4
done
echo "== kernel_bench --peer"; timeout 200 build/kernel_bench 6571 --peer --quick 2>&1 | grep -E "copy_raw_peer" | grep -E '"bytes":(16777216|67108864)' | head -14
PS_COPY_TMA=1 timeout 200 build/kernel_bench 6571 --peer --quick 2>&1 | grep -E "copy_raw_peer" | grep -E '"bytes":(16777216|67108864)' | head -14
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== train_multi joint fp8"; timeout 300 $TR --master-port 29611 tests/helpers/train_multi.py joint fp8 8 2>&1 | grep -E "rank|PASS|FAIL|rror|Check" | head -12
echo "== train_multi split bf16"; timeout 300 $TR --master-port 29612 tests/helpers/train_multi.py split bf16 8 2>&1 | grep -E "rank|PASS|FAIL|rror|Check" | head -12
echo "== bench N=2 split"; timeout 300 $TR --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench2.err | tail -c 1600; tail -3 gpurun_out/bench2.err
echo "== bench N=2 split 64MB"; timeout 300 $TR --master-port 29614 bench.py --gpus 2 --steps 10 --warmup 3 --len 67108864 --keys-per-server 8 --no-e2e 2> gpurun_out/bench2b.err | tail -c 900
echo "== bench reference N=2"; timeout 300 $TR --master-port 29615 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 2>/dev/null | tail -c 700
