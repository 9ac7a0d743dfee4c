#!/bin/bash
# N GPUs (gpurun --gpus N): the in-process multi-device path, the NCCL path, the drop-in on every GPU, bench at N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/s8_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -q > gpurun_out/s8_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/s8_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/s8_bench_n$N.json 2> gpurun_out/s8_bench_n$N.err
echo "bench rc=$?" >> gpurun_out/s8_bench_n$N.err
(cd /tmp && for d in 1 all; do echo "FRT_DEVICES=$d"; /usr/bin/time -f "wall %e s" env FRT_SKIP_PPM=1 FRT_DEVICES=$d /root/repo/oracle/_ref/cornell_shipped_b200 | grep FRT_; done) > gpurun_out/s8_dropin.txt 2>&1
tail -8 gpurun_out/s8_pytest.txt | cut -c1-250; tail -8 gpurun_out/s8_bench_n$N.err; cat gpurun_out/s8_dropin.txt
