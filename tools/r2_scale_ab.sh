#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-8}
for m in gather reduce; do
FRT_BENCH_GATHER=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 8 --warmup 3 --no-configs > /tmp/b_$m.json 2> /tmp/b_$m.err
python -c "
import json; d=json.load(open('/tmp/b_$m.json')); print('$m N=$N ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['frame_ms'], d['parity']['within_1lsb'])"
done
