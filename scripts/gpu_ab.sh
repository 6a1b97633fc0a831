#!/bin/bash
# A/B of alternative builds of the library against the in-tree one, in one gpurun call:
#   (here)  cd unbiased-inverse-volume-rendering_b200/csrc && UIVR_OUT=libuivr_x.so UIVR_NVCC_EXTRA="-D..." bash build.sh
#   gpurun -- 'bash scripts/gpu_ab.sh libuivr_x [libuivr_y ...]'
# For every alternative: the randomized parity sweep (bit-exact vs the oracle) and the config-3 timing.
mkdir -p gpurun_out
C=$PWD/unbiased-inverse-volume-rendering_b200/csrc
for l in libuivr "$@"; do
  echo "== $l"
  if [ "$l" != libuivr ]; then
    ( UIVR_LIB=$C/$l.so timeout 40 python -m pytest tests/test_gpu_parity.py -x -q -k "randomized_parity_sweep" ) 2>&1 | tail -1
  fi
  UIVR_LIB=$C/$l.so timeout 40 python scripts/quick_bench.py variant=3 reps=3 counters=0 2>&1 | tail -2
done | tee gpurun_out/ab.log
