#!/bin/bash
# Build tuning variants of libuivr.so (HERE, on CPU) into build/sweep/, one per knob setting:
#   scripts/sweep_pool.sh build name1:"-DUIVR_POOL_QUANTUM=16" name2:"..."
# and time them on the GPU box:   scripts/sweep_pool.sh run [quick_bench args]
set -e
cd "$(dirname "$0")/.."
mkdir -p sweep
if [ "$1" = build ]; then
  shift
  rm -f sweep/*.so
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    ( UIVR_OUT=$PWD/sweep/libuivr_$name.so UIVR_NVCC_EXTRA="$flags" bash unbiased-inverse-volume-rendering_b200/csrc/build.sh && echo "built $name [$flags]" ) &
  done
  wait
else
  shift
  for so in sweep/*.so; do
    echo "== $so"
    UIVR_LIB=$PWD/$so timeout 200 python scripts/quick_bench.py variant=${SWEEP_VARIANT:-3} reps=3 "$@" 2>&1 | grep "Msamples" | tail -1
  done
fi
