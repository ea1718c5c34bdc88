#!/bin/bash
# copy engine in isolation over NVLink: worker CTAs x chunk size
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for c in 149 297; do for k in 32 64 128 256; do
  echo "== engine_bench --peer ctas=$c chunk=${k}KB"
  PS_ENGINE_CHUNK_KB=$k timeout 120 build/engine_bench --peer --ctas $c 2>&1 | grep '"engine"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['bytes']>=65536: print('   ', d['bytes'], round(d['us_per_msg'],2),'us', d['GBps'],'GB/s')"
done; done
