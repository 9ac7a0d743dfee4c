#!/bin/bash
cd "$(dirname "$0")/.."
for g in 2 4 8 16 32; do echo "light group $g"; FRT_LIGHT_GROUP=$g python tools/ncu_frame.py 3 | tail -1 | sed 's/.*light_final/light_final/' | cut -c1-60; done
for b in 24 48 96; do echo "shadow blocks x$b"; FRT_SHADOW_BLOCKS=$((148*b)) python tools/ncu_frame.py 3 | tail -1 | cut -c1-60; done
