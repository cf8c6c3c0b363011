#!/bin/bash
# round 2, call v: PRMT mask expansion in the dz builders, selective mbarrier hint: parity tests, timings, ncu of fused fwd + kNN
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_tc.py tests/test_gpu_tf32x3.py -q -x > gpurun_out/v_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/v_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16,tf32x3 train > gpurun_out/v_mode_train.log 2>&1; grep -v Warn gpurun_out/v_mode_train.log | head -34
timeout 300 python scripts/r02/mode_times.py mixed16,f16+unf infer > gpurun_out/v_mode_infer.log 2>&1; grep -v Warn gpurun_out/v_mode_infer.log | grep -E "==|fused|hidden|pair_kernel|knn|device time"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_f16_pair_agg_fused" --launch-skip 5 --launch-count 1 \
   -o gpurun_out/v_fused_fwd -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/v_ncu1.log 2>&1; echo "ncu fused exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_table_split --launch-skip 3 --launch-count 1 \
   -o gpurun_out/v_knn_split -f python scripts/knn_probe.py 512 > gpurun_out/v_ncu_knn.log 2>&1; echo "ncu knn exit $?"
