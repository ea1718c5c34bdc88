#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1 PS_BENCH_WATCHDOG_S=250
echo "== multi-GPU tests"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -n 12
echo "== kernel_bench update flavours (auto selection)"
timeout 200 build/kernel_bench 6571 --quick 2>&1 | grep -i "update\|quant\|fp8" | cut -c1-220
echo "== llama-1b N=2: ours (joint, fp8 wire) / ours symmetric / nccl-ddp"
for extra in "" "--symmetric" "--symmetric --nvls-reduce --grad-wire bf16" "--impl nccl-ddp"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29932 \
    bench.py --metric llama --model llama-1b --seq-len 4096 --gpus 2 --steps 6 --warmup 3 --no-e2e $extra 2>>gpurun_out/mid_l2.err \
    | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('[$extra]', round(d['value']), 'tok/s', round(d['ms_per_step'],1),'ms', 'mfu', d.get('mfu_vs_sustained_bf16') and round(d['mfu_vs_sustained_bf16'],3), d.get('server'))
    except Exception: pass"
done
tail -n 5 gpurun_out/mid_l2.err | cut -c1-300
