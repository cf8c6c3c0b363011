#!/bin/bash
# round 2, call ao: L2 bulk prefetch of the next tile's gather rows by the TMA warp of the fused EdgeConv forward
mkdir -p gpurun_out
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ao_mode_train.log 2>&1; grep -v Warn gpurun_out/ao_mode_train.log | head -5
timeout 300 python scripts/r02/fused_roles.py > gpurun_out/ao_fused_roles.log 2>&1; grep -E "full" gpurun_out/ao_fused_roles.log
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_train_step.py -q -x > gpurun_out/ao_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/ao_pytest.log
