#!/bin/bash
# copy engine in isolation: LDG/STG workers vs TMA workers, local and over NVLink
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for t in 0 1; do for k in 128 256 512; do
  for p in "" "--peer"; do
  echo "== engine_bench $p tma=$t chunk=${k}KB"
  PS_ENGINE_TMA=$t PS_ENGINE_CHUNK_KB=$k timeout 120 build/engine_bench $p 2>&1 | grep '"engine"\|MISMATCH\|failures\|timeout\|error' | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'pattern' in d: print('   ', d['pattern'], d['bytes'], d['GBps_per_direction'],'GB/s per direction')
    elif 'bytes' in d:
        if d['bytes'] in (1024, 1048576, 4096000, 16777216): print('   ', d['bytes'], round(d['us_per_msg'],2),'us', d['GBps'],'GB/s')
    else: print('   ', d)"
done; done; done
