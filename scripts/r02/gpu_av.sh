#!/bin/bash
# round 2, call av: plain epilogue without the local-memory copy of the accumulator chunk; both weight-gradient kernels on one wave
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tf32x3.py tests/test_gpu_train_step.py -q -x > gpurun_out/av_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/av_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/av_mode_train.log 2>&1; grep -v Warn gpurun_out/av_mode_train.log | head -14
timeout 300 python scripts/r02/mode_times.py f16 infer > gpurun_out/av_mode_infer.log 2>&1; grep -v Warn gpurun_out/av_mode_infer.log | head -8
