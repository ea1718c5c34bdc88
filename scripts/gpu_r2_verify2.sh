#!/bin/bash
# Round 2, multi-GPU verification of HEAD (run with gpurun --gpus 2/4/8): multi-GPU tests incl. the
# previously gated ones, then bench.py at every power of two up to the GPU count.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
NG=$(nvidia-smi -L | wc -l)
echo "== GPUs: $NG"
nvidia-smi topo -m 2>/dev/null | head -n 12
echo "== multi-GPU tests (gated ones on)"
PSLITE_TEST_NVLS_REDUCE=1 PSLITE_TEST_NCCL_VAN=1 timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -n 25
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  echo "== bench pushpull N=$n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29930+n)) \
    bench.py --gpus $n --steps 20 --warmup 3 ${BENCH_EXTRA} 2>gpurun_out/r2_b$n.err | tee gpurun_out/r2_bench$n.json | tail -c 1500
  echo
  tail -n 3 gpurun_out/r2_b$n.err
done
