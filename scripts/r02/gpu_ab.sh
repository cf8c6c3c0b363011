#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -x -k "fused or mixed16" > gpurun_out/ab_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ab_pytest.log
timeout 300 python scripts/r02/fused_roles.py > gpurun_out/ab_fused_roles.log 2>&1; grep -v Warn gpurun_out/ab_fused_roles.log | head -8
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ab_mode_train.log 2>&1; grep -v Warn gpurun_out/ab_mode_train.log | head -5
