#!/bin/bash
# round 2, call o: launch list of the mixed16 training step + ncu --set full of its dominant kind::f16 kernels
mkdir -p gpurun_out
timeout 300 python scripts/r02/train_only.py mixed16 3 > gpurun_out/o_train_only.log 2>&1; echo "train_only exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/o_launches_train_mixed16.csv \
   python scripts/r02/train_only.py mixed16 3 > gpurun_out/o_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 300 python scripts/r02/profile_f16_kernels.py > gpurun_out/o_profile_plain.log 2>&1; echo "plain exit $?"; tail -1 gpurun_out/o_profile_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf_pair_dual_scatter|gemm_tc_pair_kernel|gemm_bf_wgrad" \
   --launch-skip 6 --launch-count 3 -o gpurun_out/o_f16_kernels -f python scripts/r02/profile_f16_kernels.py > gpurun_out/o_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/o_ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edge_hidden_fwd_node_bf16|edge_mask_bwd_bf16" --launch-skip 5 --launch-count 2 \
   -o gpurun_out/o_edge_producers -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/o_ncu_prod.log 2>&1
echo "ncu prod exit $?"; tail -2 gpurun_out/o_ncu_prod.log
