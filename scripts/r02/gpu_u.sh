#!/bin/bash
# round 2, call u: kNN shared cut + mbarrier suspend-time hint: parity tests, timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_bf16.py tests/test_gpu_tc.py -q -x > gpurun_out/u_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/u_pytest.log
timeout 300 python scripts/knn_probe.py 512 > gpurun_out/u_knn_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/u_knn_probe.log
timeout 300 python scripts/r02/mode_times.py mixed16,mixed16+unf,tf32x3 train > gpurun_out/u_mode_train.log 2>&1; grep -v Warn gpurun_out/u_mode_train.log
timeout 300 python scripts/r02/mode_times.py mixed16,f16+unf infer > gpurun_out/u_mode_infer.log 2>&1; grep -v Warn gpurun_out/u_mode_infer.log | grep -E "==|fused|hidden|pair_kernel|knn|device time"
