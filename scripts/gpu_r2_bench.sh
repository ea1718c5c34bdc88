#!/bin/bash
# bench.py (the driver's command lines) at every power of two up to the GPU count, ours then reference
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1 PS_BENCH_WATCHDOG_S=200
NG=$(nvidia-smi -L | wc -l)
show() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'value' not in d: print(d); continue
    print('$1', 'N=%d'%d['n_gpus'], d.get('impl'), round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],3), 'e2e', d.get('e2e') and round(d['e2e']['value'],1), 'launches', d.get('gpu_launches'), [(s['msg_bytes'], round(s['GBps'],1), round(s['us_per_key'],2)) for s in d.get('sweep',[])], d.get('clocks'))"; }
for n in ${ONLY_N:-1 2 4 8}; do
  [ $n -le $NG ] || continue
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29950+n))"; fi
  timeout 400 $L bench.py --gpus $n --steps 20 --warmup 3 ${BENCH_EXTRA} 2>gpurun_out/bb_$n.err | tee gpurun_out/bench_ours_n$n.json | show ours
  tail -n 2 gpurun_out/bb_$n.err | cut -c1-300
  if [ -z "$SKIP_REF" ]; then
    timeout 400 $L bench.py --impl reference --gpus $n --steps 20 --warmup 3 2>gpurun_out/br_$n.err | tee gpurun_out/bench_ref_n$n.json | show ref
  fi
done
