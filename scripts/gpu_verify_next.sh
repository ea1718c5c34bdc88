#!/bin/bash
# First GPU call of the next session: everything that was written after the GPU budget of round 1
# ran out, cheapest first, each step under its own timeout. Run with `gpurun --gpus 2`.
#   1. the regular GPU test-suite + smoke (1 GPU)
#   2. bit-exactness of the push kernels against their CPU twins
#   3. 2-GPU training tests incl. NVLS multicast fan-out, then the gated ones:
#      in-switch reduction (multimem.ld_reduce) and the nccl van
#   4. bench.py at N=1 and N=2 (the multi-peer descriptor-batch fix)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
echo "== 1. pytest -m gpu (without the multi-GPU module); PSLITE_TEST_UNVERIFIED=1 adds the checkpoint round trip"
export PSLITE_TEST_UNVERIFIED=1
timeout 900 python -m pytest tests -m gpu -x -q --ignore=tests/test_multigpu.py 2>&1 | tail -n 8
echo "== 2. kernels written after the GPU budget ran out (multi-segment copy, host/GPU bit-exactness)"
PSLITE_TEST_UNVERIFIED=1 timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_kernels_host.py -m gpu -q -k "multi_segment or wire_bytes or tiny_and_ragged" 2>&1 | tail -n 8
echo "== 2b. two-group update kernel (PS_UPDATE_X2=1): same numerics tests, then its bandwidth"
PS_UPDATE_X2=1 timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "fused_adamw or fused_sgd" 2>&1 | tail -n 4
for x2 in 0 1; do PS_UPDATE_X2=$x2 timeout 200 build/kernel_bench 2>/dev/null | grep -i update | head -6 | sed "s/^/x2=$x2 /"; done
echo "== 2c. TMA-staged update kernel (PS_UPDATE_TMA=1): numerics first (under timeout: a wrong mbarrier phase would hang), then bandwidth"
PS_UPDATE_TMA=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "fused_adamw or fused_sgd" 2>&1 | tail -n 4
for t in 0 1; do PS_UPDATE_TMA=$t timeout 200 build/kernel_bench 2>/dev/null | grep -i update | head -6 | sed "s/^/tma=$t /"; done
echo "== 3a. multi-GPU module, verified flows"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -n 8
echo "== 3b. in-switch gradient reduction"
PSLITE_TEST_NVLS_REDUCE=1 timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -k in_switch 2>&1 | tail -n 15
echo "== 3c. nccl van"
PSLITE_TEST_NCCL_VAN=1 timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -k nccl 2>&1 | tail -n 15
echo "== 4. bench N=1"
timeout 400 python bench.py --steps 20 --warmup 3 2>gpurun_out/v_b1.err | tee gpurun_out/v_bench1.json | tail -c 900
echo "== 4-. bench N=1 with the thread structure of the measured runs (no inline dispatch, fixed short poll windows)"
PS_SERVER_INLINE=0 PS_WORKER_INLINE=0 PS_SPIN_MAX_US=0 timeout 400 python bench.py --steps 20 --warmup 3 --no-e2e 2>gpurun_out/v_b1q.err | tee gpurun_out/v_bench1_queued.json | tail -c 500
echo "== 4a. PCIe bound of the end-to-end pass, then the pass through the native staged call"
timeout 120 python scripts/pcie_probe.py | tee gpurun_out/v_pcie.json
timeout 400 python bench.py --steps 20 --warmup 3 --e2e-native 2>gpurun_out/v_b1n.err | tee gpurun_out/v_bench1_native.json | tail -c 700
echo "== 4. bench N=1 with launch coalescing"
PS_COALESCE_LAUNCHES=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-e2e 2>gpurun_out/v_b1c.err | tee gpurun_out/v_bench1_coalesce.json | tail -c 600
echo "== 4. bench N=1 with the fused push-pull operation reported as well"
timeout 400 python bench.py --steps 20 --warmup 3 --no-e2e --fused-pushpull 2>gpurun_out/v_b1f.err | tee gpurun_out/v_bench1_fused.json | tail -c 700
echo "== 4. bench N=2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29931 \
  bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/v_b2.err | tee gpurun_out/v_bench2.json | tail -c 900
echo "== 4. llama-1b N=2 joint, unicast / multicast / in-switch reduce"
for extra in "" "--lazy-wait" "--fused-pushpull" "--symmetric" "--symmetric --nvls-reduce --grad-wire bf16"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29932 \
    bench.py --metric llama --model llama-1b --seq-len 4096 --gpus 2 --steps 4 --warmup 2 --no-e2e $extra 2>>gpurun_out/v_l2.err \
    | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('$extra', round(d['value']), 'tok/s', d.get('server'))
    except Exception: pass"
done
