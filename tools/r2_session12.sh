#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l > gpurun_out/s12_ngpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/s12_bench_n$N.json 2> gpurun_out/s12_bench_n$N.err
echo "bench rc=$?" >> gpurun_out/s12_bench_n$N.err
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/s12_pytest.txt 2>&1
(cd /tmp && for d in 1 all; do echo "FRT_DEVICES=$d"; s=$(date +%s.%N); FRT_SKIP_PPM=1 FRT_DEVICES=$d /root/repo/oracle/_ref/cornell_shipped_b200 | grep FRT_B200; e=$(date +%s.%N); echo "program wall $(echo "$e - $s" | bc) s"; done) > gpurun_out/s12_dropin.txt 2>&1
tail -3 gpurun_out/s12_pytest.txt; tail -4 gpurun_out/s12_bench_n$N.err; cat gpurun_out/s12_dropin.txt
