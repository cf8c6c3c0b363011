#!/bin/bash
# round 2, call r: A/B of the fused / two-kernel forward (training + inference), ncu --set full of the three builder kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -x > gpurun_out/r_pytest_bf16.log 2>&1; echo "pytest bf16 exit $?"; tail -3 gpurun_out/r_pytest_bf16.log
timeout 300 python scripts/r02/mode_times.py mixed16,mixed16+unf,mixed16+unf+dz train > gpurun_out/r_mode_train.log 2>&1; grep -v Warn gpurun_out/r_mode_train.log | grep -E "==|fused|hidden|pair_kernel<3>|pair_agg"
timeout 300 python scripts/r02/mode_times.py mixed16,mixed16+unf,f16,f16+unf infer > gpurun_out/r_mode_infer.log 2>&1; grep -v Warn gpurun_out/r_mode_infer.log | grep -E "==|fused|hidden|pair_kernel<3>|pair_agg"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_f16_pair_agg_fused" --launch-skip 5 --launch-count 1 \
   -o gpurun_out/r_fused_fwd -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/r_ncu1.log 2>&1; echo "ncu fused exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scatter_build|wgrad_build" --launch-skip 2 --launch-count 2 \
   -o gpurun_out/r_bwd_builders -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/r_ncu2.log 2>&1; echo "ncu bwd exit $?"
ls -la gpurun_out/*.ncu-rep
