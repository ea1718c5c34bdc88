#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== pytest new kernels + trainer"; timeout 600 python -m pytest tests -m gpu -x -q -k "fused or trainer or smoke" 2>&1 | tail -6
echo "== llama-1b attn backends (N=1, seq 4096, no ckpt)"
for ab in auto cudnn flash; do
  timeout 300 python bench.py --metric llama --model llama-1b --seq-len 4096 --steps 5 --warmup 2 --no-e2e --ckpt-layers 0 --attn-backend $ab 2>gpurun_out/ab_$ab.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('$ab', round(d['value']), 'tok/s mfu', round(d['mfu_vs_sustained_bf16'],3), 'mem', d.get('peak_torch_mem_gb'))
    except Exception as e: pass"
  grep -E "Error|error" gpurun_out/ab_$ab.err | head -2
done
for ck in 8 0; do
  echo "== llama3-8b N=1 fused ops ckpt=$ck"
  timeout 600 python bench.py --metric llama --steps 3 --warmup 2 --ckpt-layers $ck --no-e2e 2>gpurun_out/l8f_$ck.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(round(d['value']), 'tok/s', round(d['ms_per_step'],1),'ms mfu', round(d['mfu_vs_sustained_bf16'],3), 'mem', d.get('peak_torch_mem_gb'))
    except Exception as e: pass"
  grep -E "OutOfMemory|out of memory" gpurun_out/l8f_$ck.err | head -2
done
