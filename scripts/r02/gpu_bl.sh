#!/bin/bash
# round 2, call bl: builders of the fused forward prefetch the gather pieces of K block kb + 2 into L1 (no registers held)
mkdir -p gpurun_out
GNB_FUSED_PF=2 timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_slots8.py -q -x -k "fused" > gpurun_out/bl_pytest.log 2>&1; echo "pytest exit $?"; tail -1 gpurun_out/bl_pytest.log
for m in 0 1; do
GNB_FUSED_PF=$m timeout 300 python scripts/r02/mode_times.py f16 infer > gpurun_out/bl_mode_infer_pf$m.log 2>&1; grep -v Warn gpurun_out/bl_mode_infer_pf$m.log | grep "==\|agg_fused"
done
for m in 0 2; do
GNB_FUSED_PF=$m timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bl_mode_train_pf$m.log 2>&1; grep -v Warn gpurun_out/bl_mode_train_pf$m.log | grep "==\|agg_fused"
done
