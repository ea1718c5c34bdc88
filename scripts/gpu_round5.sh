#!/bin/bash
# 2-GPU: NVLS/symmetric-memory pull fan-out validation, updated bench numbers, ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== train_multi joint fp8 SYMM"; timeout 300 $TR --master-port 29811 tests/helpers/train_multi.py joint fp8 8 symm 2>&1 | grep -E "rank|PASS|FAIL|rror|Check|multimem|illegal" | head -12
echo "== train_multi split bf16 SYMM"; timeout 300 $TR --master-port 29812 tests/helpers/train_multi.py split bf16 8 symm 2>&1 | grep -E "rank|PASS|FAIL|rror|Check|illegal" | head -12
echo "== kernel_bench (new update math)"; CUDA_VISIBLE_DEVICES=0 timeout 300 build/kernel_bench 6571 --quick 2>&1 | grep -E "check|update_adamw|kernel_launches" | head -20
echo "== pytest gpu kernels"; CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -3
echo "== bench N=1 (new host path)"; CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 3 --sweep 1024,1048576,67108864 2>gpurun_out/b1.err | tail -c 1800; tail -2 gpurun_out/b1.err
echo "== bench N=2 split"; timeout 300 $TR --master-port 29813 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e 2>gpurun_out/b2.err | tail -c 900
echo "== llama-1b N=2 joint unicast vs symmetric"
timeout 400 $TR --master-port 29814 bench.py --metric llama --model llama-1b --seq-len 4096 --steps 6 --warmup 2 --no-e2e --ckpt-layers 0 2>gpurun_out/l2a.err | tail -c 700
timeout 400 $TR --master-port 29815 bench.py --metric llama --model llama-1b --seq-len 4096 --steps 6 --warmup 2 --no-e2e --ckpt-layers 0 --symmetric 2>gpurun_out/l2b.err | tail -c 700; tail -3 gpurun_out/l2b.err | cut -c1-300
echo "== ncu captures (1 GPU)"
export CUDA_VISIBLE_DEVICES=0
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_update -s 1 -c 1 -o gpurun_out/prof_update_bf16_w1 -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/ncu1.err; echo rc=$?
timeout 300 $NCU -k regex:k_update -s 145 -c 1 -o gpurun_out/prof_update_fp8_w4_fan5 -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/ncu2.err; echo rc=$?
timeout 300 $NCU -k regex:k_copy_vec16 -s 420 -c 1 -o gpurun_out/prof_copy_ldg -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/ncu3.err; echo rc=$?
PS_COPY_TMA=1 timeout 300 $NCU -k regex:k_copy_tma -s 420 -c 1 -o gpurun_out/prof_copy_tma -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/ncu4.err; echo rc=$?
timeout 300 $NCU -k regex:k_quant_fp8_block -s 5 -c 1 -o gpurun_out/prof_quant_fp8 -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/ncu5.err; echo rc=$?
ls -la gpurun_out/*.ncu-rep 2>/dev/null; tail -2 gpurun_out/ncu1.err
