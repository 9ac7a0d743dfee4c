#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "6 4" "4 2" "3 1" "8 6" "12 8"; do set -- $cfg; echo "RUN_MIN=$1 RUN_LEAF=$2";
  FRT_RUN_MIN=$1 FRT_RUN_LEAF=$2 python tools/dragons_perf.py 2>&1 | sed -n 2,3p | cut -c1-110
  FRT_RUN_MIN=$1 FRT_RUN_LEAF=$2 python tools/sibenik_perf.py 400 500 4 2>&1 | sed -n 2,3p | cut -c1-110
done
