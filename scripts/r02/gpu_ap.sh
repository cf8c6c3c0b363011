#!/bin/bash
# round 2, call ap: fused EdgeConv forward builders refill every row's gather registers right after the row is built
mkdir -p gpurun_out
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ap_mode_train.log 2>&1; grep -v Warn gpurun_out/ap_mode_train.log | head -5
timeout 300 python scripts/r02/fused_roles.py > gpurun_out/ap_fused_roles.log 2>&1; grep -E "full" gpurun_out/ap_fused_roles.log
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_train_step.py -q -x > gpurun_out/ap_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/ap_pytest.log
