#!/bin/bash
# round 2, call s: new kNN scan (16-byte loads + FIFO insertion) and leaner fused-forward builders: parity tests, timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_bf16.py -q -x > gpurun_out/s_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/s_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16,mixed16+unf train > gpurun_out/s_mode_train.log 2>&1; grep -v Warn gpurun_out/s_mode_train.log | grep -E "==|fused|hidden|pair_kernel<3>|knn|device time"
timeout 300 python scripts/r02/mode_times.py mixed16,f16,f16+unf infer > gpurun_out/s_mode_infer.log 2>&1; grep -v Warn gpurun_out/s_mode_infer.log | grep -E "==|fused|hidden|pair_kernel<3>|knn|device time"
