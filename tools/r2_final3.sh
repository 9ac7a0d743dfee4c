#!/bin/bash
# Final code, one B200 (the bench line follows in tools/r2_final1.sh once the traffic file is committed): pytest -m gpu, smoke, launch list of the bench frame, k_shadow_f32 capture + traffic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/rep
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.txt
python tools/ncu_frame.py 3 > gpurun_out/r2_frame_plain.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_frame.csv python tools/ncu_frame.py 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_shadow_f32 -s 1 -c 1 -o /tmp/rep/k_shadow_f32 python tools/ncu_frame.py 3 > gpurun_out/r2_ncu_k_shadow_f32.log 2>&1
python tools/ncu_summary.py /tmp/rep/k_shadow_f32.ncu-rep 0 > gpurun_out/r2_k_shadow_f32.txt 2>/dev/null
echo >> gpurun_out/r2_k_shadow_f32.txt
python tools/ncu_lines.py /tmp/rep/k_shadow_f32.ncu-rep 30 >> gpurun_out/r2_k_shadow_f32.txt 2>/dev/null
python tools/ncu_traffic.py /tmp/rep/k_shadow_f32.ncu-rep "ncu --set full --clock-control none -k regex:^k_shadow_f32 -s 1 -c 1 python tools/ncu_frame.py 3" > gpurun_out/r2_traffic_k_shadow_f32.json
python tools/stage_probe.py oracle/_ref/blobs/bounding_boxes.frt > gpurun_out/r2_dragons_plain.txt 2>&1
python tools/stage_probe.py oracle/_ref/blobs/sibenik_surrogate.frt 400 500 4 > gpurun_out/r2_sibenik_plain.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu.txt; tail -1 gpurun_out/r2_smoke.txt | cut -c1-200; cat gpurun_out/r2_traffic_k_shadow_f32.json | head -20
