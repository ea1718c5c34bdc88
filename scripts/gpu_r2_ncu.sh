#!/bin/bash
# ncu captures of the round-2 kernels (1 GPU). The copy engine is a host-fed persistent kernel and cannot
# be replayed by ncu; its evidence is engine_bench + the SASS listing. Captured here: the TMA-staged fused
# update (default for bf16 slots), the LDG update with fp8 slots, the fp8 quantiser, the raw copy + signal tail.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_update_tma -s 1 -c 1 -o gpurun_out/r2_update_tma_bf16 -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/ncu1.err; echo "update_tma rc=$?"
PS_UPDATE_TMA=0 timeout 300 $NCU -k regex:k_update -s 8 -c 1 -o gpurun_out/r2_update_ldg_fp8 -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/ncu2.err; echo "update_ldg rc=$?"
timeout 300 $NCU -k regex:k_quant_fp8_block -s 1 -c 1 -o gpurun_out/r2_quant_fp8 -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/ncu3.err; echo "quant rc=$?"
echo "== launch list of the flagship step (bench.py N=1, 2 steps)"
PSLITE_NO_AUTOBUILD=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
tail -n 3 gpurun_out/ncu_bench.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
