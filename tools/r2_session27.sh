#!/bin/bash
# N = 2: the rows pushed into rank 0's canvas (CUDA IPC) against the NCCL gather; the two-GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -4
for mode in push gather; do
  FRT_BENCH_GATHER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-configs > gpurun_out/s27_n2_$mode.json 2> gpurun_out/s27_n2_$mode.err
  echo "$mode rc=$?"; tail -2 gpurun_out/s27_n2_$mode.err | cut -c1-300
  python -c "
import json; d=json.loads(open('gpurun_out/s27_n2_$mode.json').read().strip().splitlines()[-1]); print('$mode N=2 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['frame_ms'], d['parity'])"
done
