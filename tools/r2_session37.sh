#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-configs > gpurun_out/s37_n2.json 2> gpurun_out/s37_n2.err
echo "bench rc=$?"; tail -1 gpurun_out/s37_n2.err | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/s37_n2.json').read().strip().splitlines()[-1]); print('N=2 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['frame_ms'], d['parity']['within_1lsb'], d['config']['rows_to_rank0'][:40])"
