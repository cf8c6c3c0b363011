#!/bin/bash
# round 2, call an: lane-interleaved PQ layout for the fused EdgeConv forward (half the L1 wavefronts of the builders' gathers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_train_step.py tests/test_gpu_dynedge.py tests/test_gpu_config0.py -q -x > gpurun_out/an_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/an_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/an_mode_train.log 2>&1; grep -v Warn gpurun_out/an_mode_train.log | head -6
GNB_PQ_NATURAL=1 timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/an_mode_train_natural.log 2>&1; grep -v Warn gpurun_out/an_mode_train_natural.log | head -6
timeout 300 python scripts/r02/fused_roles.py > gpurun_out/an_fused_roles.log 2>&1; grep -E "full" gpurun_out/an_fused_roles.log
