#!/bin/bash
# round 2, call ae: bench (all legs) on the final kernels, launch list of the mixed16 step, ncu --set full of the step's
# dominant kernels and of the HBM-bound tail kernels (pooling, global variables, kNN)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/ae_bench_1gpu.json 2> gpurun_out/ae_bench_1gpu.err; echo "bench exit $?"; tail -2 gpurun_out/ae_bench_1gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/ae_bench_1gpu.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d["inference"]["value"], d["inference"]["precision"][:30], "launches", d.get("gpu_launches_per_step"))
r = d["roofline"]; print(r["kernel"][:50], r["launch_ms"], r["achieved"], r["frac"], r["executed_frac"])
ri = d["inference"]["roofline"]; print("inf", ri["kernel"][:50], ri["launch_ms"], ri["achieved"], ri["frac"])
for k in d["kernels"]["kernels"][:14]: print(k)
PY
timeout 300 python scripts/r02/train_only.py mixed16 3 > gpurun_out/ae_train_only.log 2>&1; echo "train_only exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ae_launches_train_mixed16.csv \
   python scripts/r02/train_only.py mixed16 3 > gpurun_out/ae_ncu_list.log 2>&1; echo "ncu list exit $?"
python scripts/summarize_launches.py gpurun_out/ae_launches_train_mixed16.csv 30 > gpurun_out/ae_launches_train_mixed16_summary.txt 2>&1; head -24 gpurun_out/ae_launches_train_mixed16_summary.txt
for k in gemm_f16_pair_agg_fused gemm_f16_pair_scatter_build gemm_f16_wgrad_build; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 5 --launch-count 1 \
     -o gpurun_out/ae_$k -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/ae_ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"segment_pool|global_vars|edge_dz_prep|knn_table_split" --launch-skip 11 --launch-count 11 \
   -o gpurun_out/ae_tail_kernels -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/ae_ncu_tail.log 2>&1; echo "ncu tail exit $?"
ls -la gpurun_out/ae_*.ncu-rep
