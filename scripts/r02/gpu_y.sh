#!/bin/bash
# round 2, call y: programmatic dependent launch on every kernel: full GPU test suite, timings with and without the attribute
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/y_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -4 gpurun_out/y_pytest_gpu.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/y_mode_train.log 2>&1; grep -v Warn gpurun_out/y_mode_train.log | head -6
GNB_PDL=0 timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/y_mode_train_nopdl.log 2>&1; grep -v Warn gpurun_out/y_mode_train_nopdl.log | head -3
timeout 300 python scripts/r02/mode_times.py f16,mixed16 infer > gpurun_out/y_mode_infer.log 2>&1; grep -v Warn gpurun_out/y_mode_infer.log | grep "=="
GNB_PDL=0 timeout 300 python scripts/r02/mode_times.py f16 infer 2>&1 | grep "=="
timeout 300 python scripts/r02/gaps.py mixed16 2>&1 | grep -E "^kernels|^gap"
