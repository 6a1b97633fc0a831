#!/bin/bash
# Last gpurun call of a round: parity tests first, then smoke, then the bench line (bounded).
mkdir -p gpurun_out
( time timeout ${PYTEST_LIMIT:-420} python -m pytest tests -m gpu -x -q -rP ) > gpurun_out/pytest_gpu.log 2>&1
grep -a "losses:" gpurun_out/pytest_gpu.log | tail -3; tail -6 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
