#!/bin/bash
# rows copied in front of the frame's final wait, host barrier in PushGather, k_shadow_mesh with one frame: tests + N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s33_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s33_pytest.txt
tail -3 gpurun_out/s33_pytest.txt | cut -c1-200
python tools/dragons_perf.py 2>&1 | sed -n 2p | cut -c1-70
python tools/sibenik_perf.py 400 500 4 2>&1 | sed -n 2p | cut -c1-70
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-configs > gpurun_out/s33_n2.json 2> gpurun_out/s33_n2.err
echo "bench rc=$?"; tail -2 gpurun_out/s33_n2.err | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/s33_n2.json').read().strip().splitlines()[-1]); print('N=2 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['frame_ms'], d['parity'])"
