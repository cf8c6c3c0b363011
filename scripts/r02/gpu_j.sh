#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -s ${1:+-k "$1"} > gpurun_out/j_pytest_bf16.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|error|assert|timeout" gpurun_out/j_pytest_bf16.log | cut -c1-300 | head -30
timeout 600 python scripts/r02/mode_times.py ${2:-mixed16} train > gpurun_out/k_modes_train.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_train.log | grep -v -i Warn
timeout 600 python scripts/r02/mode_times.py ${3:-f16} infer > gpurun_out/k_modes_infer.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_infer.log | grep -v -i Warn
