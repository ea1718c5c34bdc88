#!/bin/bash
# Second GPU contact: pytest -m gpu, smoke, nvl test_benchmark, bench.py (1 GPU)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
for len in 1024 1048576 67108864; do
  echo "== test_benchmark nvl 1w1s same GPU len=$len"
  PS_VAN_TYPE=nvl TEST_NUM_GPU_WORKER=1 TEST_NUM_GPU_SERVER=1 PS_CUDA_DEVICE=0 NUM_KEY_PER_SERVER=8 \
    TOTAL_DURATION=40 LOG_DURATION=20 BENCH_JSON=1 timeout 120 scripts/local.sh 1 1 build/test_benchmark $len 10 1 \
    > gpurun_out/tb_nvl_$len.log 2>&1; echo "rc=$?"
  grep -E "goodput|Check failed|rror" gpurun_out/tb_nvl_$len.log | head -6
done
echo "== bench.py ours N=1"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_ours_1.json; tail -5 gpurun_out/bench_ours_1.err
echo "== bench.py reference N=1"; timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref_1.json 2> gpurun_out/bench_ref_1.err; echo "rc=$?"; tail -c 1200 gpurun_out/bench_ref_1.json
echo "== bench.py llama-1b quick"; timeout 600 python bench.py --metric llama --model llama-1b --seq-len 4096 --steps 4 --warmup 2 > gpurun_out/bench_llama1b.json 2> gpurun_out/bench_llama1b.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_llama1b.json; tail -8 gpurun_out/bench_llama1b.err
