#!/bin/bash
# NVLink: LDG/STG vs TMA copy kernels, one direction vs both directions at once
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for t in 0 1; do
  echo "== kernel_bench --peer PS_COPY_TMA=$t"
  PS_COPY_TMA=$t timeout 200 build/kernel_bench --peer --quick 2>&1 | grep copy_raw_peer | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ', d['bytes'], 'ctas', d['ctas'], round(d['us'],2),'us', d['algo_GBps'],'GB/s', d['flavour'])"
  echo "== kernel_bench local PS_COPY_TMA=$t"
  PS_COPY_TMA=$t timeout 200 build/kernel_bench --quick 2>&1 | grep '"copy_raw"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['ctas'] in (0,148,296): print('   ', d['bytes'], 'ctas', d['ctas'], round(d['us'],2),'us', d['algo_GBps'],'GB/s r+w', d['flavour'])"
done
echo "== engine_bench --peer (bidirectional at the end)"
timeout 200 build/engine_bench --peer 2>&1 | grep pattern
