#!/bin/bash
# quick GPU iteration: parity tests + fwd/bwd timing of config 3
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
python scripts/quick_bench.py variant=0 reps=3 > gpurun_out/quick_v0.log 2>&1; tail -6 gpurun_out/quick_v0.log
