#!/bin/bash
# 1 GPU: full gpu test-suite + smoke, Llama-3-8B memory/ckpt exploration, pushpull N=1.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench N=1"; timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/b1.err | tail -c 1200; tail -2 gpurun_out/b1.err
for ck in 16 8; do
  echo "== llama3-8b N=1 ckpt_layers=$ck"
  timeout 600 python bench.py --metric llama --steps 3 --warmup 2 --ckpt-layers $ck --no-e2e > gpurun_out/l8_ck$ck.json 2> gpurun_out/l8_ck$ck.err; echo "rc=$?"
  tail -c 700 gpurun_out/l8_ck$ck.json; grep -E "OutOfMemory|out of memory|Error" gpurun_out/l8_ck$ck.err | head -3
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
