#!/bin/bash
# quick GPU iteration: parity tests + config-3 timing of one variant
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-2}; do
  timeout 300 python scripts/quick_bench.py variant=$v reps=3 > gpurun_out/quick_v$v.log 2>&1; tail -6 gpurun_out/quick_v$v.log
done
