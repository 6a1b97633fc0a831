#!/bin/bash
# One gpurun call of the tuning loop: GPU parity suite on the in-tree build, config-3 timing of the in-tree build
# and of every alternative build under sweep/ (scripts/sweep_pool.sh build ...), then one ncu capture of a step.
#   TAG=r2c bash scripts/gpu_sweep.sh
TAG=${TAG:-sweep}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
tail -8 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python scripts/quick_bench.py variant=3 reps=3 > gpurun_out/quick_default_$TAG.log 2>&1; tail -4 gpurun_out/quick_default_$TAG.log
SWEEP_VARIANT=3 timeout 900 bash scripts/sweep_pool.sh run counters=0 2>&1 | tee gpurun_out/sweep_$TAG.log
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pool -s 3 -c 3 -f -o gpurun_out/prof_$TAG \
    python scripts/profile_step.py variant=3 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
fi
ls -la gpurun_out | tail -12
