#!/bin/bash
# round 2, call bf: the zeroing of the scatter target and of the bias scratch folded into the dz prep launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_slots8.py tests/test_gpu_train_step.py tests/test_gpu_dynedge.py -q -x > gpurun_out/bf_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/bf_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bf_mode_train.log 2>&1; grep -v Warn gpurun_out/bf_mode_train.log | head -16
