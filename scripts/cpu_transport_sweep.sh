#!/bin/bash
# CPU-only transport comparison on one host (no GPU needed): the reference's ZMQ van (ipc://)
# against this repo's TCP van (unix sockets + shared-memory rings) and one-sided shm van, with
# the reference's own test_benchmark protocol: 1 worker + 1 server, 40 keys, push+pull per key.
# usage: scripts/cpu_transport_sweep.sh [out.jsonl]
cd "$(dirname "$0")/.."
OUT=${1:-profiles/cpu_transport_sweep.jsonl}
: > "$OUT"
REF=baseline/_ref/bin/test_benchmark
# rounds per log window / in total: enough of them that small messages are not a startup measurement
rounds() { if [ "$1" -le 262144 ]; then echo "TOTAL_DURATION=1500 LOG_DURATION=500"; else echo "TOTAL_DURATION=30 LOG_DURATION=10"; fi; }
run_ours() {  # van len extra-env...
  local van=$1 len=$2; shift 2
  env "$@" PS_VAN_TYPE=$van DMLC_LOCAL=1 NUM_KEY_PER_SERVER=40 $(rounds $len) \
    timeout 300 scripts/local.sh 1 1 build/test_benchmark $len 100000 1 2>&1 | grep -oE "goodput: [0-9.e+-]+" | tail -n 1 | cut -d' ' -f2
}
run_ref() {
  local len=$1 port=$((12000 + RANDOM % 20000))
  local common="DMLC_NUM_WORKER=1 DMLC_NUM_SERVER=1 DMLC_PS_ROOT_URI=127.0.0.1 DMLC_PS_ROOT_PORT=$port DMLC_NODE_HOST=127.0.0.1 DMLC_GROUP_SIZE=1 DMLC_LOCAL=1 NUM_KEY_PER_SERVER=40 $(rounds $len)"
  export LD_LIBRARY_PATH=baseline/_ref/lib:$LD_LIBRARY_PATH
  env $common DMLC_ROLE=scheduler timeout 300 $REF $len 10 1 >/dev/null 2>&1 &
  env $common DMLC_ROLE=server timeout 300 $REF $len 10 1 >/dev/null 2>&1 &
  env $common DMLC_ROLE=worker timeout 300 $REF $len 10 1 2>&1 | grep -oE "goodput: [0-9.e+-]+" | tail -n 1 | cut -d' ' -f2
  wait
}
for len in 1024 16384 262144 1048576 4096000 16777216 67108864; do
  r=$(run_ref $len)
  t=$(run_ours tcp $len)
  tn=$(run_ours tcp $len PS_SHM_PIPE=0)
  s=$(run_ours shm $len TEST_EXPORTABLE_VALS=1)
  sq=$(run_ours shm $len TEST_EXPORTABLE_VALS=1 BENCHMARK_INLINE=0)
  echo "{\"msg_bytes\": $len, \"reference_zmq_ipc_gbps\": ${r:-null}, \"tcp_van_gbps\": ${t:-null}, \"tcp_van_no_rings_gbps\": ${tn:-null}, \"shm_onesided_van_gbps\": ${s:-null}, \"shm_onesided_van_queued_dispatch_gbps\": ${sq:-null}}" | tee -a "$OUT"
done
