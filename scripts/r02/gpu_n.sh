#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_tc.py tests/test_gpu_tf32x3.py -q -m gpu -x -s > gpurun_out/n_pytest.log 2>&1; echo "pytest exit $?"; grep -E "^(f16|mixed16) rel errors|passed|failed|Error" gpurun_out/n_pytest.log | cut -c1-300
timeout 600 python scripts/r02/mode_times.py ${1:-mixed16,f16} train > gpurun_out/k_modes_train.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_train.log | grep -v -i Warn
timeout 600 python scripts/r02/mode_times.py ${2:-tf32,f16,bf16} infer > gpurun_out/k_modes_infer.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_infer.log | grep -v -i Warn
