#!/bin/bash
# round 2, GPU call L: parity suite, A/B of the north_star variants (scatter aggregation / v4, SMEM walk table via TMA),
# bench.py with the new keys, launch list + ncu capture of one step
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r2l.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2l.log
timeout 900 bash scripts/gpu_ab.sh libuivr_nsmall libuivr_smemtab libuivr_match libuivr_v4 libuivr_matchv4 > gpurun_out/ab_r2l.log 2>&1; cat gpurun_out/ab_r2l.log
timeout 900 python bench.py > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; cut -c1-400 gpurun_out/bench_r2l.json; tail -3 gpurun_out/bench_r2l.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2l.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r2l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pool -s 3 -c 3 -f -o gpurun_out/prof_r2l \
    python scripts/profile_step.py variant=3 > gpurun_out/ncu_full_r2l.log 2>&1
tail -3 gpurun_out/ncu_full_r2l.log
ls -la gpurun_out | tail -8
