#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu exit $?"
cat gpurun_out/bench_2gpu.json | cut -c1-300; tail -5 gpurun_out/bench_2gpu.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "1gpu exit $?"
cat gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
cat gpurun_out/bench_ref.json | cut -c1-400
