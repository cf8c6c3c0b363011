#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -s ${1:+-k "$1"} > gpurun_out/j_pytest_bf16.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|error|assert" gpurun_out/j_pytest_bf16.log | cut -c1-300 | head -30
timeout 300 python scripts/r02/profile_masked.py > gpurun_out/p_masked_plain.log 2>&1; grep -v Warn gpurun_out/p_masked_plain.log; timeout 600 python scripts/r02/mode_times.py mixed16 train > gpurun_out/k_modes_train.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_train.log | grep -v -i Warn
