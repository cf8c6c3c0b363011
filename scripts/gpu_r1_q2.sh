#!/bin/bash
# build o: tests, bench, kNN probe, launch list and ncu --set full of the two dominant launches
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['inference'], d['gpu_launches'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['forward_launch']['launch_ms'], d['cpu_baseline'])"

timeout 300 python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-inference > gpurun_out/bench_short_tf32.json 2> gpurun_out/bench_short_tf32.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_tf32.csv \
    python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-inference > gpurun_out/ncu_tf32.log 2>&1
echo "ncu launch list exit $?"
timeout 300 python scripts/profile_top_kernel.py > gpurun_out/top_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair --launch-skip 4 --launch-count 2 \
   -o gpurun_out/q_gemm_tc_pair -f python scripts/profile_top_kernel.py > gpurun_out/ncu_top.log 2>&1
echo "ncu top exit $?"; tail -3 gpurun_out/ncu_top.log
