#!/bin/bash
# round 2, call ar: 8-slot edge layout (device-side switch) in the four per-edge kernels of the mixed16 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slots8.py -q -x > gpurun_out/ar_pytest_slots8.log 2>&1; echo "slots8 exit $?"; tail -15 gpurun_out/ar_pytest_slots8.log
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_train_step.py -q -x > gpurun_out/ar_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ar_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ar_mode_train.log 2>&1; grep -v Warn gpurun_out/ar_mode_train.log | head -22
GNB_SLOTS9=1 timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ar_mode_train_slots9.log 2>&1; grep -v Warn gpurun_out/ar_mode_train_slots9.log | head -22
