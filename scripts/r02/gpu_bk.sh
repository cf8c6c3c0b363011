#!/bin/bash
# round 2, call bk: edge-slot flag on many CTAs (word zeroed with the scale words)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slots8.py tests/test_gpu_bf16.py tests/test_gpu_train_step.py tests/test_gpu_dynedge.py -q -x > gpurun_out/bk_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/bk_pytest.log
timeout 300 python scripts/r02/mode_times.py f16 infer > gpurun_out/bk_mode_infer.log 2>&1; grep -v Warn gpurun_out/bk_mode_infer.log | grep "==\|slot_flag\|device"
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bk_mode_train.log 2>&1; grep -v Warn gpurun_out/bk_mode_train.log | grep "==\|slot_flag\|device"
