#!/bin/bash
# round 2, call bd: fused EdgeConv forward walking its tiles in descending order (the tail of PQ is what L2 still holds)
mkdir -p gpurun_out
GNB_FUSED_REV=1 timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_slots8.py -q -x -k "fused or training_step" > gpurun_out/bd_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/bd_pytest.log
for r in 0 1 0 1; do
GNB_FUSED_REV=$r timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bd_mode_train_rev$r.log 2>&1; grep -v Warn gpurun_out/bd_mode_train_rev$r.log | grep "==\|agg_fused"
done
GNB_FUSED_REV=1 timeout 300 python scripts/r02/mode_times.py f16 infer > gpurun_out/bd_mode_infer_rev1.log 2>&1; grep -v Warn gpurun_out/bd_mode_infer_rev1.log | grep "==\|agg_fused"
