#!/bin/bash
# bench (validates the JSON line), then one ncu --set full capture of the two dominant launches
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['inference']['value']); print(json.dumps(d['roofline'])[:1500])"
timeout 300 python scripts/profile_top_kernel.py > gpurun_out/top_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair --launch-skip 4 --launch-count 2 \
   -o gpurun_out/k_gemm_tc_pair -f python scripts/profile_top_kernel.py > gpurun_out/ncu_top.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_top.log
