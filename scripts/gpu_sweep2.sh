#!/bin/bash
# One gpurun call: every alternative build under sweep/ (scripts/sweep_pool.sh build ...) passes the randomized parity
# sweep and is timed on config 3, with the in-tree build first and last (drift of the box during the call).
#   TAG=r2o bash scripts/gpu_sweep2.sh
TAG=${TAG:-sweep2}
mkdir -p gpurun_out
{
echo "== in-tree"; timeout 120 python scripts/quick_bench.py variant=3 reps=3 counters=0 2>&1 | grep Msamples | tail -1
for so in sweep/*.so; do
  echo "== $so"
  ( UIVR_LIB=$PWD/$so timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "randomized_parity_sweep" ) 2>&1 | tail -1
  UIVR_LIB=$PWD/$so timeout 120 python scripts/quick_bench.py variant=3 reps=3 counters=0 2>&1 | grep Msamples | tail -1
done
echo "== in-tree"; timeout 120 python scripts/quick_bench.py variant=3 reps=3 counters=0 2>&1 | grep Msamples | tail -1
} | tee gpurun_out/sweep2_$TAG.log
