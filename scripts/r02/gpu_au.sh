#!/bin/bash
# round 2, call au: node-level weight-gradient kernel with 1 / 2 waves of split-K CTAs; ncu of the PQ GEMM (split tf32 forward)
mkdir -p gpurun_out
for w in 2 1; do
GNB_WG_WAVES=$w timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/au_mode_train_w$w.log 2>&1; grep -v Warn gpurun_out/au_mode_train_w$w.log | grep "==\|wgrad\|device time"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel --launch-skip 16 --launch-count 1 \
   -o gpurun_out/au_pq_gemm -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/au_ncu_pq.log 2>&1; echo "ncu exit $?"
