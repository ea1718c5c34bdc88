#!/bin/bash
# 8-GPU sanity + scaling: pushpull N=4, N=8 (split), llama joint N=8.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== train_multi joint fp8 N=8"; timeout 300 $TR --nproc-per-node 8 --master-port 29711 tests/helpers/train_multi.py joint fp8 6 2>&1 | grep -E "rank 0|rank 7|PASS|FAIL|rror|Check" | head -8
echo "== bench N=8 split (4w+4s)"; timeout 400 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --steps 20 --warmup 3 --sweep 1024,65536,1048576,16777216,67108864,268435456 2> gpurun_out/bench8.err > gpurun_out/bench8.json; echo rc=$?; tail -c 2500 gpurun_out/bench8.json; tail -3 gpurun_out/bench8.err | cut -c1-300
echo "== bench N=4 split (2w+2s)"; timeout 300 $TR --nproc-per-node 4 --master-port 29713 bench.py --gpus 4 --steps 20 --warmup 3 2> gpurun_out/bench4.err > gpurun_out/bench4.json; echo rc=$?; tail -c 1200 gpurun_out/bench4.json
echo "== reference N=8"; timeout 300 $TR --nproc-per-node 8 --master-port 29714 bench.py --impl reference --gpus 8 --steps 20 --warmup 3 2>/dev/null | tail -c 500
echo "== llama3-8b joint N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29715 bench.py --metric llama --gpus 8 --steps 4 --warmup 2 --ckpt-layers 8 2> gpurun_out/llama8_n8.err > gpurun_out/llama8_n8.json; echo rc=$?; tail -c 1500 gpurun_out/llama8_n8.json; tail -3 gpurun_out/llama8_n8.err | cut -c1-300
echo "== llama3-8b split N=8 (4w+4s)"; timeout 600 $TR --nproc-per-node 8 --master-port 29716 bench.py --metric llama --gpus 8 --topology split --steps 4 --warmup 2 --ckpt-layers 0 --no-e2e 2> gpurun_out/llama8_split.err > gpurun_out/llama8_split.json; echo rc=$?; tail -c 1000 gpurun_out/llama8_split.json; tail -3 gpurun_out/llama8_split.err | cut -c1-300
