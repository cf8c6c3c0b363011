#!/bin/bash
# round 2, final call (1 GPU): the driver's GPU test command, smoke(), bench (all legs + reference arm + workloads), launch list of
# the mixed16 step, ncu --set full of the step's dominant kernels and of the PQ GEMM after the local-store fix
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/ay_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ay_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ay_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/ay_smoke.log
timeout 900 python bench.py > gpurun_out/ay_bench_1gpu.json 2> gpurun_out/ay_bench_1gpu.err; echo "bench exit $?"; tail -2 gpurun_out/ay_bench_1gpu.err
timeout 900 python bench.py --impl reference > gpurun_out/ay_bench_reference.json 2> gpurun_out/ay_bench_reference.err; echo "reference exit $?"
for w in highmult20k percentile16 prometheus50 microbench; do
  timeout 600 python bench.py --workload $w > gpurun_out/ay_bench_$w.json 2> gpurun_out/ay_bench_$w.err; echo "$w exit $?"
done
timeout 600 python bench.py --workload tito256 --precision tf32x3 > gpurun_out/ay_bench_tito256.json 2> gpurun_out/ay_bench_tito256.err; echo "tito256 exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ay_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d["inference"]["value"], "launches", d.get("gpu_launches_per_step"))
r = d["roofline"]; print(r["bound"], r["kernel"][:50], r["launch_ms"], r["achieved"], r["peak"], r["frac"], r.get("tensor_view"))
for k in d["kernels"]["kernels"][:14]: print(k)
PY
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ay_mode_train.log 2>&1; grep -v Warn gpurun_out/ay_mode_train.log | head -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ay_launches_train_mixed16.csv \
   python scripts/r02/train_only.py mixed16 3 > gpurun_out/ay_ncu_list.log 2>&1; echo "ncu list exit $?"
python scripts/summarize_launches.py gpurun_out/ay_launches_train_mixed16.csv 30 > gpurun_out/ay_launches_train_mixed16_summary.txt 2>&1; head -8 gpurun_out/ay_launches_train_mixed16_summary.txt
for k in gemm_f16_pair_agg_fused gemm_f16_pair_scatter_build gemm_f16_wgrad_build; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 5 --launch-count 1 \
     -o gpurun_out/ay_$k -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/ay_ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel --launch-skip 16 --launch-count 1 \
   -o gpurun_out/ay_pq_gemm_split_fwd -f python scripts/r02/train_only.py mixed16 2 > gpurun_out/ay_ncu_pq.log 2>&1; echo "ncu pq exit $?"
ls -la gpurun_out/ay_*.ncu-rep
