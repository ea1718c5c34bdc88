#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for p in "" "--peer"; do
  echo "== engine_bench $p (defaults: TMA workers, 512 KB chunks, chunks taken dynamically)"
  timeout 120 build/engine_bench $p 2>&1 | grep '"engine"\|MISMATCH\|failures\|timeout\|error' | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'pattern' in d: print('   ', d['pattern'], d['bytes'], d['GBps_per_direction'],'GB/s', d.get('us_per_msg',''))
    elif 'bytes' in d: print('   ', d['bytes'], round(d['us_per_msg'],2),'us', d['GBps'],'GB/s', 'one-at-a-time', d['one_at_a_time_us'])
    else: print('   ', d)"
done
export PSLITE_NO_AUTOBUILD=1
SKIP_REF=1 bash scripts/gpu_r2_bench.sh
