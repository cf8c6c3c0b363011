#!/bin/bash
# round 2, call ba: dz prep kernel writes the row-major mask through shared memory (16-byte pieces)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_slots8.py tests/test_gpu_train_step.py -q -x > gpurun_out/ba_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/ba_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ba_mode_train.log 2>&1; grep -v Warn gpurun_out/ba_mode_train.log | grep "==\|dz_prep\|device time"
