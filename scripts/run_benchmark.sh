#!/bin/bash
# Multi-port TCP benchmark (parity: reference tests/run_benchmark.sh:15-22):
# 40,960,000 B x 25 rounds, 10 keys per server, DMLC_NUM_PORTS rails through the multivan.
cd "$(dirname "$0")/.."
export PS_VAN_TYPE=multivan DMLC_NUM_PORTS=${DMLC_NUM_PORTS:-2} NUM_KEY_PER_SERVER=10
export TOTAL_DURATION=${TOTAL_DURATION:-25} LOG_DURATION=${LOG_DURATION:-5}
exec scripts/local.sh ${NUM_SERVERS:-1} ${NUM_WORKERS:-1} build/test_benchmark 40960000 25 1
