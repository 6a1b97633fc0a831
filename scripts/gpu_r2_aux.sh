#!/bin/bash
# Round-2 timings of the BASELINE configurations that are not the bench metric, the production ray-batch shape and the
# nerf integrator, after the full GPU suite.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_aux.log 2>&1; tail -5 gpurun_out/pytest_gpu_aux.log
timeout 600 python scripts/config_bench.py > gpurun_out/r02_config_bench.jsonl 2> gpurun_out/config_bench.err; cat gpurun_out/r02_config_bench.jsonl | cut -c1-400; tail -2 gpurun_out/config_bench.err
timeout 300 python scripts/batch_bench.py > gpurun_out/r02_batch_bench.txt 2>&1; tail -5 gpurun_out/r02_batch_bench.txt
timeout 300 python scripts/nerf_bench.py > gpurun_out/r02_nerf_bench.json 2> gpurun_out/nerf_bench.err; cut -c1-600 gpurun_out/r02_nerf_bench.json
