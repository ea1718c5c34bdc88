#!/bin/bash
# Gather / scatter / dense-reduce stress over joint nodes (parity: reference test_stress.sh:2,46):
# 30,000,000 B values, 8 sessions per node. NODES joint processes on this host.
cd "$(dirname "$0")/.."
export JOINT=1 BENCHMARK_NTHREAD=${BENCHMARK_NTHREAD:-8} DEBUG_MODE=1
N=${NODES:-2}
exec scripts/local.sh $N $N build/test_benchmark_stress ${LEN:-30000000} ${REPEAT:-10}
