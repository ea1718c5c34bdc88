#!/bin/bash
# Llama-3-8B PS training on ONE B200 (worker + server co-located): does it fit, how fast.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
export PYTORCH_CUDA_ALLOC_CONF=${PYTORCH_CUDA_ALLOC_CONF:-}
for ck in 32 24; do
  echo "== llama3-8b seq 8192 ckpt_layers=$ck"
  nvidia-smi --query-gpu=memory.used --format=csv,noheader -lms 2000 > gpurun_out/mem_$ck.txt &
  SMI_PID=$!   # stopped by its own pid below, never by name
  timeout 900 python bench.py --metric llama --model llama3-8b --seq-len 8192 --steps 3 --warmup 2 --ckpt-layers $ck \
     > gpurun_out/llama8b_ck$ck.json 2> gpurun_out/llama8b_ck$ck.err; echo "rc=$?"
  tail -c 1400 gpurun_out/llama8b_ck$ck.json; tail -4 gpurun_out/llama8b_ck$ck.err | cut -c1-300
  sort -n gpurun_out/mem_$ck.txt | tail -1
  kill "$SMI_PID" 2>/dev/null; wait "$SMI_PID" 2>/dev/null
done
