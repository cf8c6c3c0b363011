#!/bin/bash
# round 2, call i: full GPU test suite + smoke + default bench on the tree as of session 2 start
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/i_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/i_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/i_smoke.log
timeout 600 python bench.py > gpurun_out/i_bench_1gpu.json 2> gpurun_out/i_bench_1gpu.err; echo "bench exit $?"; head -c 500 gpurun_out/i_bench_1gpu.json
