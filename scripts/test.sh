#!/bin/bash
# Push/pull preset of the reference's two-node script (test.sh:2,21-27,59-77: 4,096,000 B, mode 1).
#   scripts/test.sh local                      one box, TCP van
#   scripts/test.sh nvl                        one box, values in HBM over the NVLink van (1 GPU per process)
#   scripts/test.sh remote <role> <root_ip>    two boxes: run `remote scheduler <ip>` and `remote server <ip>` on
#                                              the first, `remote worker <ip>` on the second (<ip> = address of
#                                              the first box). PS_VAN_TYPE=nvl with TEST_NUM_GPU_WORKER/SERVER=1
#                                              keeps the values in HBM: the vans stage them through the host,
#                                              because the peer sits on another machine (docs/how_to.md).
cd "$(dirname "$0")/.."
mode=${1:-local}
export BENCHMARK_NTHREAD=${BENCHMARK_NTHREAD:-1} TOTAL_DURATION=${TOTAL_DURATION:-100} LOG_DURATION=${LOG_DURATION:-20}
if [ "$mode" = "remote" ]; then
  role=${2:?usage: scripts/test.sh remote scheduler|server|worker <root_ip>}
  export DMLC_PS_ROOT_URI=${3:?root ip missing} DMLC_PS_ROOT_PORT=${DMLC_PS_ROOT_PORT:-8123}
  export DMLC_NUM_SERVER=${NUM_SERVERS:-1} DMLC_NUM_WORKER=${NUM_WORKERS:-1} DMLC_ROLE=$role
  # DMLC_INTERFACE / DMLC_NODE_HOST pick the address this process advertises (default: first non-loopback)
  exec build/test_benchmark 4096000 100000 1
fi
if [ "$mode" = "nvl" ]; then
  export PS_VAN_TYPE=nvl TEST_NUM_GPU_WORKER=1 TEST_NUM_GPU_SERVER=1 WORKER_GPU_BASE=0 SERVER_GPU_BASE=${NUM_WORKERS:-1}
fi
exec scripts/local.sh ${NUM_SERVERS:-1} ${NUM_WORKERS:-1} build/test_benchmark 4096000 100000 1
