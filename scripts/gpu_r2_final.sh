#!/bin/bash
# Round-2 evidence call (one B200): full GPU parity suite, smoke(), the bench line of both arms, and the ncu launch list of
# the bench command (per-launch times are cold-cache and serialised: only the kernels' SHARES of a step are comparable).
TAG=${TAG:-r2final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_$TAG.txt 2>&1; nproc >> gpurun_out/gpu_$TAG.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
if [ -n "$WITH_NCU_FULL" ]; then
# one config-3 step under `ncu --set full`: the constants bench.py needs for roofline.issue / traffic, stamped with the
# fingerprint of the kernel sources of THIS build
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pool -s 3 -c 3 -f -o gpurun_out/prof_$TAG \
    python scripts/profile_step.py variant=3 > gpurun_out/ncu_full_$TAG.log 2>&1
python scripts/ncu_constants.py gpurun_out/prof_$TAG.ncu-rep "r02 gpurun call $TAG (ncu --set full --clock-control none, scripts/profile_step.py variant=3, launches 4-6)" \
    && cp profiles/kernel_constants.json gpurun_out/kernel_constants_$TAG.json
fi
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
if [ -z "$NO_REF" ]; then
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cat gpurun_out/bench_ref_$TAG.json
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu_$TAG.log 2>&1
tail -2 gpurun_out/bench_under_ncu_$TAG.log | cut -c1-300
