#!/bin/bash
# round 2, call m: full suite with mixed16 in the parity tests, smoke, default bench (mixed16)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/m_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/m_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/m_smoke.log
timeout 900 python bench.py > gpurun_out/m_bench_1gpu.json 2> gpurun_out/m_bench_1gpu.err; echo "bench exit $?"; head -c 1500 gpurun_out/m_bench_1gpu.json; tail -5 gpurun_out/m_bench_1gpu.err
