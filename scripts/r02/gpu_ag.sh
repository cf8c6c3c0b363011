#!/bin/bash
# round 2, call ag: fused forward builders (ReLU on the conversion, packed mask, own mapping for the 16-channel tail K block)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_train_step.py tests/test_gpu_dynedge.py -q -x > gpurun_out/ag_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ag_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ag_mode_train.log 2>&1; grep -v Warn gpurun_out/ag_mode_train.log | head -14
timeout 300 python scripts/r02/fused_roles.py > gpurun_out/ag_fused_roles.log 2>&1; grep -E "full|wgrad|dgrad|agg" gpurun_out/ag_fused_roles.log
