#!/bin/bash
# First GPU contact: kernel self-checks + microbench, then the nvl van end to end.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpus.txt 2>&1
echo "== kernel_bench (ldg)"; timeout 300 build/kernel_bench 6571 --quick > gpurun_out/kernel_bench_ldg.jsonl 2> gpurun_out/kernel_bench_ldg.err; echo "rc=$?"; tail -3 gpurun_out/kernel_bench_ldg.err
echo "== kernel_bench (tma)"; PS_COPY_TMA=1 timeout 300 build/kernel_bench 6571 --quick > gpurun_out/kernel_bench_tma.jsonl 2> gpurun_out/kernel_bench_tma.err; echo "rc=$?"; tail -3 gpurun_out/kernel_bench_tma.err
grep -E "check|copy_raw\"" gpurun_out/kernel_bench_ldg.jsonl | head -30
grep -E "check|copy_raw\"" gpurun_out/kernel_bench_tma.jsonl | head -30
grep -E "push_|update" gpurun_out/kernel_bench_ldg.jsonl
for len in 1024 1048576 67108864; do
  echo "== test_benchmark nvl 1w1s same GPU len=$len"
  PS_VAN_TYPE=nvl TEST_NUM_GPU_WORKER=1 TEST_NUM_GPU_SERVER=1 PS_CUDA_DEVICE=0 NUM_KEY_PER_SERVER=8 \
    TOTAL_DURATION=40 LOG_DURATION=20 BENCH_JSON=1 timeout 120 scripts/local.sh 1 1 build/test_benchmark $len 10 1 \
    > gpurun_out/tb_nvl_$len.log 2>&1; echo "rc=$?"
  grep -E "goodput|Check failed|rror" gpurun_out/tb_nvl_$len.log | head -6
done
