#!/bin/bash
# Multi-GPU evidence: bench.py at N GPUs of one box, launched the way the driver does; plus the gloo/NCCL sharding tests.
N=${N:-2}; TAG=${TAG:-r2n$N}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/gpus_$TAG.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json | cut -c1-1500; tail -5 gpurun_out/bench_$TAG.err
if [ -n "$WITH_REF" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
cat gpurun_out/bench_ref_$TAG.json | cut -c1-600
fi
