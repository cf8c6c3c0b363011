#!/bin/bash
# round 2, call ac: one weight-packing launch per step: full GPU test suite + timings
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/ac_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -4 gpurun_out/ac_pytest_gpu.log
timeout 300 python scripts/r02/mode_times.py mixed16,tf32x3,bf16 train > gpurun_out/ac_mode_train.log 2>&1; grep -v Warn gpurun_out/ac_mode_train.log | grep -E "==|device time"
timeout 300 python scripts/r02/mode_times.py f16,mixed16 infer > gpurun_out/ac_mode_infer.log 2>&1; grep -v Warn gpurun_out/ac_mode_infer.log | grep -E "==|device time"
