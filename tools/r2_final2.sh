#!/bin/bash
# Evidence of the round's final code on one B200 (session 2): pytest -m gpu, smoke, bench lines of both arms, ncu launch
# list of the bench frame, full captures of k_shadow_f32 (+ DRAM traffic), of the mesh kernels (C4 stand-in) and of the
# photon-map kernels (C5), summarised on the box (tools/ncu_summary.py, ncu_lines.py, ncu_traffic.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/rep
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n1.err
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
python tools/sibenik_perf.py 400 500 4 > gpurun_out/r2_sibenik_plain.txt 2>&1
python tools/dragons_perf.py > gpurun_out/r2_dragons_plain.txt 2>&1
for k in k_shadow_mesh k_extend; do
  ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -o /tmp/rep/mesh_$k python tools/sibenik_perf.py 400 500 4 > gpurun_out/r2_ncu_mesh_$k.log 2>&1
  summ mesh_$k /tmp/rep/mesh_$k.ncu-rep 0
  rm -f /tmp/rep/mesh_$k.ncu-rep
done
python tools/gi_stage_probe.py 400 > gpurun_out/r2_gi_plain.txt 2>&1
for k in k_knn_cell k_fg_trace; do
  ncu --set full --clock-control none --import-source on -k regex:^$k -s 2 -c 1 -o /tmp/rep/gi_$k python tools/gi_stage_probe.py 400 > gpurun_out/r2_ncu_gi_$k.log 2>&1
  summ gi_$k /tmp/rep/gi_$k.ncu-rep 0
  rm -f /tmp/rep/gi_$k.ncu-rep
done
tail -3 gpurun_out/r2_pytest_gpu.txt; tail -1 gpurun_out/r2_smoke.txt; tail -2 gpurun_out/r2_bench_n1.err; ls gpurun_out | grep "^r2_" | head -40
