#!/bin/bash
# launch list of a bench frame with the last commit of the round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/ncu_frame.py 3 > gpurun_out/r2d_frame_plain.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2d_launches_frame.csv python tools/ncu_frame.py 3 > /dev/null 2>&1
tail -1 gpurun_out/r2d_frame_plain.txt | cut -c1-200
