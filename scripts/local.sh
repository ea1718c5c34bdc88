#!/bin/bash
# Run a ps-lite style job on this host: 1 scheduler + S servers + W workers.
# usage: scripts/local.sh <num_servers> <num_workers> <binary> [args...]
# (parity: reference tests/local.sh:8-37). Extra env is inherited.
# JOINT=1: start num_workers joint processes instead; SET_RANKS=1: export DMLC_RANK=i. If any process
# fails, the rest are killed so a crash never turns into a hang.
set -u
if [ $# -lt 3 ]; then
  echo "usage: $0 num_servers num_workers bin [args..]"; exit 1
fi
export DMLC_NUM_SERVER=$1; shift
export DMLC_NUM_WORKER=$1; shift
bin=$1; shift
args="$@"
export DMLC_PS_ROOT_URI=${DMLC_PS_ROOT_URI:-127.0.0.1}
export DMLC_PS_ROOT_PORT=${DMLC_PS_ROOT_PORT:-$((12000 + RANDOM % 20000))}
export DMLC_NODE_HOST=${DMLC_NODE_HOST:-127.0.0.1}
pids=()
DMLC_ROLE=scheduler ${bin} ${args} &
pids+=($!)
if [ -n "${JOINT:-}" ]; then
  # co-located mode: every process is worker i + server i (DMLC_ROLE=joint)
  for ((i=0; i<${DMLC_NUM_WORKER}; ++i)); do
    if [ -n "${WORKER_GPU_BASE:-}" ]; then export PS_CUDA_DEVICE=$((WORKER_GPU_BASE + i)); fi
    if [ "${BYTEPS_ENABLE_MIXED_MODE:-0}" != "0" ]; then
      DMLC_ROLE=joint ${bin} ${args} &   # mixed mode orders the ranks itself (plain servers first)
    else
      DMLC_ROLE=joint DMLC_RANK=$i ${bin} ${args} &
    fi
    pids+=($!)
  done
  # mixed mode (BYTEPS_ENABLE_MIXED_MODE): servers beyond the co-located ones are plain processes
  for ((i=${DMLC_NUM_WORKER}; i<${DMLC_NUM_SERVER}; ++i)); do
    if [ -n "${SERVER_GPU_BASE:-}" ]; then export PS_CUDA_DEVICE=$((SERVER_GPU_BASE + i - DMLC_NUM_WORKER)); fi
    DMLC_ROLE=server ${bin} ${args} &
    pids+=($!)
  done
else
  for ((i=0; i<${DMLC_NUM_SERVER}; ++i)); do
    if [ -n "${SERVER_GPU_BASE:-}" ]; then export PS_CUDA_DEVICE=$((SERVER_GPU_BASE + i)); fi
    if [ -n "${SET_RANKS:-}" ]; then export DMLC_RANK=$i; fi
    DMLC_ROLE=server ${bin} ${args} &
    pids+=($!)
  done
  for ((i=0; i<${DMLC_NUM_WORKER}; ++i)); do
    if [ -n "${WORKER_GPU_BASE:-}" ]; then export PS_CUDA_DEVICE=$((WORKER_GPU_BASE + i)); fi
    if [ -n "${SET_RANKS:-}" ]; then export DMLC_RANK=$i; fi
    DMLC_ROLE=worker ${bin} ${args} &
    pids+=($!)
  done
fi
# fail fast: stop everybody as soon as one process fails. `wait -n` returns 127 once no child is
# left to report — it can get there early, because bash may retire several children that exited
# together on one call — so the authoritative exit codes are collected per pid afterwards
# (`wait <pid>` also answers for children that are already gone).
rc=0
remaining=${#pids[@]}
while [ $remaining -gt 0 ]; do
  wait -n; st=$?
  remaining=$((remaining - 1))
  if [ $st -eq 127 ]; then break; fi
  if [ $st -ne 0 ]; then
    rc=$st
    for p in "${pids[@]}"; do kill $p 2>/dev/null; done
    break
  fi
done
for p in "${pids[@]}"; do
  wait $p 2>/dev/null; st=$?
  if [ $rc -eq 0 ] && [ $st -ne 0 ] && [ $st -ne 127 ]; then rc=$st; fi
done
exit $rc
