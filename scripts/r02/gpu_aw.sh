#!/bin/bash
# round 2, call aw: fused EdgeConv forward with plane 0 of W2 resident in shared memory
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_slots8.py tests/test_gpu_train_step.py -q -x > gpurun_out/aw_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/aw_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/aw_mode_train.log 2>&1; grep -v Warn gpurun_out/aw_mode_train.log | head -6
GNB_FUSED_WRES=0 timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/aw_mode_train_stream.log 2>&1; grep -v Warn gpurun_out/aw_mode_train_stream.log | head -6
timeout 300 python scripts/r02/mode_times.py f16,mixed16 infer > gpurun_out/aw_mode_infer.log 2>&1; grep -v Warn gpurun_out/aw_mode_infer.log | grep "==\|fused\|device time"
GNB_F16_INFER_FUSED=1 timeout 300 python scripts/r02/mode_times.py f16 infer > gpurun_out/aw_mode_infer_fused.log 2>&1; grep -v Warn gpurun_out/aw_mode_infer_fused.log | head -8
