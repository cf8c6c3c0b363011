#!/bin/bash
# round 2, call af: vectorised pooling forward: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dynedge.py tests/test_gpu_users.py -q -x > gpurun_out/af_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/af_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/af_mode_train.log 2>&1; grep -v Warn gpurun_out/af_mode_train.log | grep -E "==|pool|device"
timeout 300 python scripts/r02/mode_times.py f16 infer > gpurun_out/af_mode_infer.log 2>&1; grep -v Warn gpurun_out/af_mode_infer.log | grep -E "==|pool|device"
