#!/bin/bash
# round 2, call h: ncu --set full of the dominant launches of the tf32x3 training step (split aggregating forward GEMM,
# dual-group scattering GEMM, weight-gradient GEMM), the other workloads, reference arm
mkdir -p gpurun_out
timeout 300 python scripts/r02/train_only.py tf32x3 2 > gpurun_out/h_train_only.log 2>&1; echo "train_only exit $?"
# the 5th launch of gemm_tc_pair_kernel<true> of a step is the aggregating GEMM of conv layer 2 (PQ1, agg1, PQ2, agg2 ...): skip
# the first step (11 launches) + 3, capture agg2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel --launch-skip 22 --launch-count 4 \
   -o gpurun_out/h_gemm_tc_pair_split_fwd -f python scripts/r02/train_only.py tf32x3 2 > gpurun_out/h_ncu_split.log 2>&1
echo "ncu split exit $?"; tail -2 gpurun_out/h_ncu_split.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dual_scatter|gemm_tc_wgrad" --launch-skip 24 --launch-count 3 \
   -o gpurun_out/h_dual_scatter_wgrad -f python scripts/r02/train_only.py tf32x3 2 > gpurun_out/h_ncu_bwd.log 2>&1
echo "ncu bwd exit $?"; tail -2 gpurun_out/h_ncu_bwd.log
for w in prometheus50 highmult20k percentile16 microbench infer1024; do
  timeout 600 python bench.py --workload $w --steps 6 --warmup 2 > gpurun_out/h_bench_$w.json 2> gpurun_out/h_bench_$w.err; echo "$w exit $?"; head -c 600 gpurun_out/h_bench_$w.json; echo; tail -3 gpurun_out/h_bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_reference.json 2> gpurun_out/h_bench_reference.err; echo "ref exit $?"; head -c 400 gpurun_out/h_bench_reference.json
