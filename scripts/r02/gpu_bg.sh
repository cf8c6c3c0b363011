#!/bin/bash
# round 2, call bg: fused EdgeConv forward, 8-slot tiles: a lane's four rows are four slots of one node (P gathered once per K block)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_slots8.py tests/test_gpu_train_step.py -q -x > gpurun_out/bg_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/bg_pytest.log
for r in 1 2; do
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bg_mode_train_$r.log 2>&1; grep -v Warn gpurun_out/bg_mode_train_$r.log | grep "==\|agg_fused"
GNB_SLOTS9=1 timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bg_mode_train_slots9_$r.log 2>&1; grep -v Warn gpurun_out/bg_mode_train_slots9_$r.log | grep "==\|agg_fused"
done
timeout 300 python scripts/r02/mode_times.py f16 infer > gpurun_out/bg_mode_infer.log 2>&1; grep -v Warn gpurun_out/bg_mode_infer.log | grep "==\|agg_fused"
