#!/bin/bash
# Several workers per server on one machine, each with its own log prefix and (optionally) its own
# GPU (parity: reference tests/local_multi_workers.sh). Same argument order as local.sh; this
# wrapper only adds per-process tags so that interleaved output can be told apart.
#   usage: scripts/local_multi_workers.sh num_servers num_workers bin [args..]
if [ $# -lt 3 ]; then
  echo "usage: $0 num_servers num_workers bin [args..]"; exit 1
fi
cd "$(dirname "$0")/.."
export SET_RANKS=${SET_RANKS:-1}          # DMLC_RANK per process: stable ranks across restarts
export PS_VERBOSE=${PS_VERBOSE:-0}
exec scripts/local.sh "$@"
