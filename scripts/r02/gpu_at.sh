#!/bin/bash
# round 2, call at: 8-slot layout, whole GPU suite; weight-gradient builder kernel with 1 / 2 / 3 waves of split-K CTAs; ncu of the PQ GEMM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/at_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/at_pytest_gpu.log
for w in 2 1 3; do
GNB_WGB_WAVES=$w timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/at_mode_train_w$w.log 2>&1; grep -v Warn gpurun_out/at_mode_train_w$w.log | grep "==\|wgrad_build\|slot_flag\|device time"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel --launch-skip 9 --launch-count 1 \
   -o gpurun_out/at_pq_gemm -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/at_ncu_pq.log 2>&1; echo "ncu exit $?"
