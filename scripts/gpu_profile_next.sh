#!/bin/bash
# ncu captures for the next round (ONE GPU; never a multi-rank command): the two-group update
# kernel next to the one-group kernel, the model kernels that have no capture yet, and a launch
# list of one bench.py step with and without launch coalescing. Reports land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
NCU="ncu --set full --clock-control none --import-source on"
echo "== k_update (one group) vs k_update_x2"
timeout 300 $NCU -k regex:k_update -s 1 -c 1 -o gpurun_out/prof2_update_bf16_w1 -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/p2_1.err; echo rc=$?
PS_UPDATE_X2=1 timeout 300 $NCU -k regex:k_update_x2 -s 1 -c 1 -o gpurun_out/prof2_update_x2_bf16_w1 -f build/kernel_bench 6571 --quick > /dev/null 2>gpurun_out/p2_2.err; echo rc=$?
echo "== model kernels (RoPE-split, SwiGLU) from the fused-op unit tests"
timeout 300 $NCU -k regex:k_rope -c 1 -o gpurun_out/prof2_rope -f python -m pytest tests/test_kernels_gpu.py -q -k rope > /dev/null 2>gpurun_out/p2_3.err; echo rc=$?
timeout 300 $NCU -k regex:k_swiglu_fwd -c 1 -o gpurun_out/prof2_swiglu -f python -m pytest tests/test_kernels_gpu.py -q -k swiglu > /dev/null 2>gpurun_out/p2_4.err; echo rc=$?
echo "== launch lists of two bench steps (numbers printed under ncu are NOT bench values)"
for co in 0 1; do
  PS_COALESCE_LAUNCHES=$co timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_coalesce$co.csv python bench.py --steps 2 --warmup 3 --no-e2e > /dev/null 2>gpurun_out/p2_l$co.err
  echo "coalesce=$co rc=$? kernels: $(grep -c k_copy gpurun_out/launches_coalesce$co.csv 2>/dev/null)"
done
for f in gpurun_out/prof2_*.ncu-rep; do
  [ -f "$f" ] && ncu -i "$f" --page raw --csv > "${f%.ncu-rep}.raw.csv" 2>/dev/null
done
ls -la gpurun_out | tail -n 15
