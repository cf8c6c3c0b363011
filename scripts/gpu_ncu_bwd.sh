#!/bin/bash
# ncu --set full of the two HBM-bound backward kernels of a conv layer (edge_mask_bwd -> dz, weight-gradient GEMM over dz + h)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_wgrad|edge_mask_bwd" --launch-skip 7 --launch-count 2 \
   -o gpurun_out/s_mask_bwd_and_wgrad -f python bench.py --steps 1 --warmup 1 --repeats 1 --no-cpu-baseline --no-inference > gpurun_out/ncu_bwd.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_bwd.log
