#!/bin/bash
cd "$(dirname "$0")/.."
FRT_KNN_SORT=1 python tools/gi_stage_probe.py 400 | tail -1 | cut -c1-400
FRT_GI_QUEUE=4000000 python tools/gi_stage_probe.py 400 | tail -1 | cut -c1-400
