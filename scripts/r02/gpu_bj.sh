#!/bin/bash
# round 2, call bj: ncu --set full of the one-plane fused EdgeConv forward in the f16 inference step + launch list of that step
mkdir -p gpurun_out
timeout 300 python scripts/r02/infer_only.py f16 2 > gpurun_out/bj_infer.log 2>&1; echo "infer exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bj_launches_infer_f16.csv \
   python scripts/r02/infer_only.py f16 3 > gpurun_out/bj_ncu_list.log 2>&1; echo "ncu list exit $?"
python scripts/summarize_launches.py gpurun_out/bj_launches_infer_f16.csv 20 > gpurun_out/bj_launches_infer_f16_summary.txt 2>&1; head -10 gpurun_out/bj_launches_infer_f16_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_pair_agg_fused --launch-skip 5 --launch-count 1 \
   -o gpurun_out/bj_fused_fwd_f16x1_inference -f python scripts/r02/infer_only.py f16 2 > gpurun_out/bj_ncu_fused.log 2>&1; echo "ncu exit $?"
