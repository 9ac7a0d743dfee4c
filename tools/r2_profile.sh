#!/bin/bash
# Final evidence of round 2 on one B200: bench lines of both arms, ncu launch list and full captures of the kernels of a
# Cornell frame, of the mesh walk (C4 stand-in) and of the photon-map kernels (C5).  The captures are summarised on the box
# (tools/ncu_summary.py, ncu_lines.py, ncu_traffic.py); only the summaries and one .ncu-rep travel back (64 MiB limit).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/rep
if [ "$1" != "nobench" ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
  echo "bench rc=$?" >> gpurun_out/r2_bench_n1.err
fi
if [ "$1" = "ref" ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
fi
summ() { # name rep [kernel index]
  python tools/ncu_summary.py $2 ${3:-0} > gpurun_out/r2_$1.txt 2>/dev/null
  echo >> gpurun_out/r2_$1.txt
  python tools/ncu_lines.py $2 30 >> gpurun_out/r2_$1.txt 2>/dev/null
}
python tools/ncu_frame.py 3 > gpurun_out/r2_frame_plain.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_frame.csv python tools/ncu_frame.py 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_shadow_f32 -s 2 -c 2 -o /tmp/rep/k_shadow_f32 python tools/ncu_frame.py 3 > gpurun_out/r2_ncu_k_shadow_f32.log 2>&1
summ k_shadow_f32 /tmp/rep/k_shadow_f32.ncu-rep 0
python tools/ncu_traffic.py /tmp/rep/k_shadow_f32.ncu-rep "ncu --set full --clock-control none -k regex:^k_shadow_f32 -s 2 -c 2 python tools/ncu_frame.py 3" > gpurun_out/r2_traffic_k_shadow_f32.json
cp /tmp/rep/k_shadow_f32.ncu-rep gpurun_out/r2_k_shadow_f32.ncu-rep
for k in k_shadow_bulk k_shadow_quad k_extend k_shade k_light_pre k_light_final k_shadow_exact; do
  ncu --set full --clock-control none --import-source on -k regex:^$k -s 2 -c 1 -o /tmp/rep/$k python tools/ncu_frame.py 3 > gpurun_out/r2_ncu_$k.log 2>&1
  summ $k /tmp/rep/$k.ncu-rep 0
  rm -f /tmp/rep/$k.ncu-rep
done
python tools/sibenik_perf.py 400 500 4 > gpurun_out/r2_sibenik_plain.txt 2>&1 &&
for k in k_shadow_mesh k_extend; do
  ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -o /tmp/rep/mesh_$k python tools/sibenik_perf.py 400 500 4 > gpurun_out/r2_ncu_mesh_$k.log 2>&1
  summ mesh_$k /tmp/rep/mesh_$k.ncu-rep 0
  rm -f /tmp/rep/mesh_$k.ncu-rep
done
python tools/gi_stage_probe.py 160 > gpurun_out/r2_gi_plain.txt 2>&1 &&
for k in k_knn k_fg_trace; do
  ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -o /tmp/rep/gi_$k python tools/gi_stage_probe.py 160 > gpurun_out/r2_ncu_gi_$k.log 2>&1
  summ gi_$k /tmp/rep/gi_$k.ncu-rep 0
  rm -f /tmp/rep/gi_$k.ncu-rep
done
du -sh gpurun_out; ls gpurun_out | head -50; tail -3 gpurun_out/r2_bench_n1.err
