#!/bin/bash
# round 2, call j: 16-bit plane kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -s ${1:+-k "$1"} > gpurun_out/j_pytest_bf16.log 2>&1; echo "pytest exit $?"; grep -E "rel errors|passed|failed|Error|error" gpurun_out/j_pytest_bf16.log | cut -c1-400; tail -30 gpurun_out/j_pytest_bf16.log | cut -c1-300
timeout 600 python scripts/r02/mode_times.py tf32x3,mixed16,bf16 train > gpurun_out/k_modes_train.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_train.log | grep -v -i Warn
