#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$N.json')); print('N=$N ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['frame_ms'], d['parity']['within_1lsb'])"
