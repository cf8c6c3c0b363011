#!/bin/bash
# round 2, call w: kNN size classes (2 / 8 lanes per query), fused-forward L2 prefetch of the next tile, act_bwd_colsum geometry,
# f16 inference on the two-kernel forward: parity tests, timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_bf16.py tests/test_gpu_ops.py tests/test_gpu_dynedge.py -q -x > gpurun_out/w_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/w_pytest.log
timeout 300 python scripts/knn_probe.py 512 > gpurun_out/w_knn_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/w_knn_probe.log
timeout 300 python scripts/knn_probe.py 1024 >> gpurun_out/w_knn_probe.log 2>&1; tail -4 gpurun_out/w_knn_probe.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/w_mode_train.log 2>&1; grep -v Warn gpurun_out/w_mode_train.log | head -20
timeout 300 python scripts/r02/mode_times.py mixed16,f16 infer > gpurun_out/w_mode_infer.log 2>&1; grep -v Warn gpurun_out/w_mode_infer.log | grep -E "==|fused|hidden|pair_kernel|knn|device time|pool|global"
