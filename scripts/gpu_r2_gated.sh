#!/bin/bash
# Round 2: gated descriptor frames (in-kernel completion signal) vs the event + completion-thread path.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PSLITE_NO_AUTOBUILD=1
export PS_BENCH_WATCHDOG_S=150
NG=$(nvidia-smi -L | wc -l)
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 12
run() {  # n, tag, env...
  local n=$1 tag=$2; shift 2
  if [ "$n" = 1 ]; then
    env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --sweep 1024,65536,1048576,16777216 2>gpurun_out/g_$tag.err | tee gpurun_out/g_$tag.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$tag', 'N=$n', round(d['value'],1),'GB/s', round(d['ms_per_step']*1000/d['config']['global_batch']*d['config']['num_workers'],2),'us/key', 'launches', d['gpu_launches'], [(s['msg_bytes'], round(s['GBps'],1), round(s['us_per_key'],2)) for s in d['sweep']], d.get('van_stats_rank0'))"
  else
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29940+n)) \
      bench.py --gpus $n --steps 20 --warmup 3 --no-e2e --sweep 1024,65536,1048576,16777216 2>gpurun_out/g_$tag.err | tee gpurun_out/g_$tag.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$tag', 'N=$n', round(d['value'],1),'GB/s', round(d['ms_per_step']*1000/d['config']['global_batch']*d['config']['num_workers'],2),'us/key', 'launches', d['gpu_launches'], [(s['msg_bytes'], round(s['GBps'],1), round(s['us_per_key'],2)) for s in d['sweep']], d.get('van_stats_rank0'))"
  fi
  tail -n 3 gpurun_out/g_$tag.err | cut -c1-300
}
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  run $n gated_n$n PS_GATED_FRAMES=1
  run $n events_n$n PS_GATED_FRAMES=0
  run $n engine_n$n PS_COPY_ENGINE=1
  run $n engine_c64_n$n PS_COPY_ENGINE=1 PS_ENGINE_CTAS=65
  run $n engine_c296_n$n PS_COPY_ENGINE=1 PS_ENGINE_CTAS=296
done
