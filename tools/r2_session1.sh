#!/bin/bash
# GPU session 1 of round 2: tests, bench, drop-in timing, allocation probe, ncu launch list + full capture of k_shadow_f32
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s1_gpus.txt 2>&1
./tools/probe_alloc > gpurun_out/s1_probe_alloc.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench rc=$?" >> gpurun_out/s1_bench.err
(cd /tmp && FRT_SKIP_PPM=1 /root/repo/oracle/_ref/cornell_shipped_b200 && FRT_SKIP_PPM=1 /root/repo/oracle/_ref/cornell_shipped_b200 && FRT_LIGHT_GEN=0 FRT_SKIP_PPM=1 /root/repo/oracle/_ref/cornell_shipped_b200) > gpurun_out/s1_dropin.txt 2>&1
python tools/ncu_frame.py 3 > gpurun_out/s1_frame_plain.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s1_launches.csv python tools/ncu_frame.py 3 > gpurun_out/s1_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shadow_f32 -s 2 -c 2 -o gpurun_out/s1_k_shadow_f32 python tools/ncu_frame.py 3 > gpurun_out/s1_ncu_full.log 2>&1
tail -3 gpurun_out/s1_pytest.txt; cat gpurun_out/s1_bench.json | head -c 1500; cat gpurun_out/s1_probe_alloc.txt
