#!/bin/bash
# A/B of an alternative build of the library (UIVR_LIB) against the in-tree one: parity sweep, then config-3 timing.
mkdir -p gpurun_out
ALT=$PWD/unbiased-inverse-volume-rendering_b200/csrc/${1:?name of the alternative .so under csrc/ (built with UIVR_OUT=... UIVR_NVCC_EXTRA=-D... bash build.sh)}
( UIVR_LIB=$ALT timeout 40 python -m pytest tests/test_gpu_parity.py -x -q -k "randomized_parity_sweep" ) > gpurun_out/alt_sweep.log 2>&1; tail -2 gpurun_out/alt_sweep.log
timeout 40 python scripts/quick_bench.py variant=3 reps=4 counters=0 > gpurun_out/alt_base.log 2>&1; tail -3 gpurun_out/alt_base.log
UIVR_LIB=$ALT timeout 40 python scripts/quick_bench.py variant=3 reps=4 counters=0 > gpurun_out/alt_alt.log 2>&1; tail -3 gpurun_out/alt_alt.log
