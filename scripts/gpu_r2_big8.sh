#!/bin/bash
# The one 8-GPU call of round 2: multi-GPU tests, config 2 / config 5 in C++, bench N=4/8, Llama-3-8B N=8 vs DDP.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1 PS_BENCH_WATCHDOG_S=200
echo "== 1. multi-GPU tests (8 GPUs visible)"
timeout 420 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "several_gpus or stress_gpu or ipc_benchmark or native_symmetric" 2>&1 | tail -n 12 | cut -c1-300
echo "== 2. config 2: apps/test_benchmark 4 worker GPUs + 4 server GPUs, push+pull sweep (engine on)"
for sz in 1024 65536 1048576 4096000 16777216 67108864; do
  rounds=600; [ $sz -gt 100000 ] && rounds=120; [ $sz -gt 20000000 ] && rounds=30
  keys=40; [ $sz -gt 20000000 ] && keys=8
  PS_VAN_TYPE=nvl TEST_NUM_GPU_WORKER=1 TEST_NUM_GPU_SERVER=1 NUM_KEY_PER_SERVER=$keys WORKER_GPU_BASE=0 SERVER_GPU_BASE=4 PS_COPY_ENGINE=1 \
    TOTAL_DURATION=$rounds LOG_DURATION=$((rounds/3)) timeout 100 scripts/local.sh 4 4 build/test_benchmark $sz 100 1 > gpurun_out/cfg2_$sz.log 2>&1
  echo "  $sz B: $(grep -a goodput gpurun_out/cfg2_$sz.log | awk '{print $6}' | sort -n | tail -n 4 | tr '\n' ' ') Gbps per worker (last windows of the 4 workers)"
  grep -a "Check failed\|what()" gpurun_out/cfg2_$sz.log | head -n 2 | cut -c1-300
done
echo "== 3. config 5: test_ipc_benchmark, 8 joint GPUs, NVLS multicast pull vs unicast pull"
for nv in 1 0; do
  PS_VAN_TYPE=nvl JOINT=1 WORKER_GPU_BASE=0 IPC_NVLS_PULL=$nv IPC_VERIFY=1 NUM_KEY_PER_SERVER=10 PS_COPY_ENGINE=1 timeout 120 scripts/local.sh 8 8 build/test_ipc_benchmark 4096000 40 > gpurun_out/cfg5_nvls$nv.log 2>&1
  grep -a "goodput" gpurun_out/cfg5_nvls$nv.log | cut -c30-330 | head -n 3
  grep -a "Check failed\|what()" gpurun_out/cfg5_nvls$nv.log | head -n 2 | cut -c1-300
done
echo "== 4. bench.py N=4, N=8"
SKIP_REF=1 ONLY_N='4 8' bash scripts/gpu_r2_bench.sh
echo "== 5. Llama-3-8B, 8 GPUs, seq 8192: PS (joint 8w+8s, fp8 wire, symmetric/NVLS pull) vs NCCL DDP"
for extra in "--symmetric" "--impl nccl-ddp"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29971 \
    bench.py --metric llama --gpus 8 --steps 5 --warmup 2 $extra 2>gpurun_out/llama8.err | tee -a gpurun_out/llama8_n8.jsonl \
    | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('[$extra]', round(d['value']), 'tok/s', round(d['ms_per_step'],1),'ms/step', 'e2e', d.get('e2e') and round(d['e2e']['value']), 'mfu', d.get('mfu_vs_sustained_bf16') and round(d['mfu_vs_sustained_bf16'],3), 'mem', d.get('peak_torch_mem_gb'), d.get('server'))
    except Exception: pass"
  tail -n 3 gpurun_out/llama8.err | cut -c1-300
done
