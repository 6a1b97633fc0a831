#!/bin/bash
# gpu_sweep2.sh + one `ncu --set full` capture of a config-3 step of the in-tree build (forward, adjoint, DRT launch)
TAG=${TAG:-sweep3}
bash scripts/gpu_sweep2.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pool -s 3 -c 3 -f -o gpurun_out/prof_$TAG \
    python scripts/profile_step.py variant=3 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
