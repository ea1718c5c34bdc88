#!/bin/bash
# Round 2, first GPU call (1 GPU): does HEAD run at all? tests, smoke, kernel A/B, bench N=1.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
echo "== 1. pytest -m gpu (no multi-GPU module)"
PSLITE_TEST_UNVERIFIED=1 timeout 900 python -m pytest tests -m gpu -x -q --ignore=tests/test_multigpu.py 2>&1 | tail -n 15
echo "== 1b. smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -n 5
echo "== 2. kernel_bench default / x2 / tma"
timeout 200 build/kernel_bench 2>&1 | tee gpurun_out/r2_kernel_bench_default.txt | tail -n 60
for v in PS_UPDATE_X2 PS_UPDATE_TMA; do
  env $v=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "fused_adamw or fused_sgd" 2>&1 | tail -n 3
  env $v=1 timeout 200 build/kernel_bench 2>/dev/null | grep -i update | sed "s/^/$v /" | tee gpurun_out/r2_kernel_bench_$v.txt
done
echo "== 3. pcie probe"
timeout 120 python scripts/pcie_probe.py | tee gpurun_out/r2_pcie.json
echo "== 4. bench N=1 variants"
timeout 400 python bench.py --steps 20 --warmup 3 2>gpurun_out/r2_b1.err | tee gpurun_out/r2_bench1.json | tail -c 1200
echo
PS_SERVER_INLINE=0 PS_WORKER_INLINE=0 PS_SPIN_MAX_US=0 timeout 400 python bench.py --steps 20 --warmup 3 --no-e2e 2>gpurun_out/r2_b1q.err | tee gpurun_out/r2_bench1_queued.json | tail -c 600
echo
PS_COALESCE_LAUNCHES=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-e2e 2>gpurun_out/r2_b1c.err | tee gpurun_out/r2_bench1_coalesce.json | tail -c 600
echo
timeout 400 python bench.py --steps 20 --warmup 3 --e2e-native 2>gpurun_out/r2_b1n.err | tee gpurun_out/r2_bench1_native.json | tail -c 900
echo
echo "== 5. reference arm N=1"
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 2>gpurun_out/r2_ref1.err | tee gpurun_out/r2_ref1.json | tail -c 600
tail -n 5 gpurun_out/r2_*.err
