#!/bin/bash
cd "$(dirname "$0")/.."
for g in 1 2 4 8; do echo "pre group $g"; FRT_PRE_GROUP=$g python tools/ncu_frame.py 3 | tail -1 | cut -c1-330; done
