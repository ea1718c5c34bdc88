#!/bin/bash
# Final-ish 8-GPU numbers: pushpull 4w+4s, Llama-3-8B joint (unicast vs NVLS multicast), split.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(json.dumps({k:d.get(k) for k in ("value","unit","ms_per_step","e2e","mfu_vs_sustained_bf16","peak_torch_mem_gb","server","sweep","gpu_launches","clocks")}))
    except Exception: pass'
echo "== bench N=8 split"; timeout 400 $TR --nproc-per-node 8 --master-port 29912 bench.py --gpus 8 --steps 20 --warmup 3 --sweep 1048576,16777216,67108864 2>gpurun_out/b8.err | tee gpurun_out/bench8_v2.json | python -c "$J"
echo "== bench N=4 split"; timeout 300 $TR --nproc-per-node 4 --master-port 29913 bench.py --gpus 4 --steps 20 --warmup 3 2>gpurun_out/b4.err | tee gpurun_out/bench4_v2.json | python -c "$J"
echo "== llama3-8b joint N=8 unicast"; timeout 600 $TR --nproc-per-node 8 --master-port 29915 bench.py --metric llama --gpus 8 --steps 5 --warmup 2 2>gpurun_out/l8j.err | tee gpurun_out/llama8_joint_v2.json | python -c "$J"; tail -2 gpurun_out/l8j.err | cut -c1-200
echo "== llama3-8b joint N=8 NVLS"; timeout 600 $TR --nproc-per-node 8 --master-port 29916 bench.py --metric llama --gpus 8 --steps 5 --warmup 2 --no-e2e --symmetric 2>gpurun_out/l8s.err | tee gpurun_out/llama8_joint_nvls.json | python -c "$J"; tail -2 gpurun_out/l8s.err | cut -c1-200
echo "== llama3-8b split N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29917 bench.py --metric llama --gpus 8 --topology split --steps 5 --warmup 2 --no-e2e --symmetric 2>gpurun_out/l8p.err | tee gpurun_out/llama8_split_nvls.json | python -c "$J"; tail -2 gpurun_out/l8p.err | cut -c1-200
