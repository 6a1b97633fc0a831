#!/bin/bash
# round 2, GPU call A: parity suite on the new walker/handler split, A/B of tuning builds, ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 120 python scripts/quick_bench.py variant=3 reps=3 > gpurun_out/quick_default.log 2>&1; tail -4 gpurun_out/quick_default.log
SWEEP_VARIANT=3 timeout 900 bash scripts/sweep_pool.sh run counters=0 2>&1 | tee gpurun_out/sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pool -s 4 -c 4 -f -o gpurun_out/prof_r2a \
    python scripts/profile_step.py variant=3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
