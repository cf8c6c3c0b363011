#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/infer_launchlist.py > gpurun_out/infer_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_infer.csv \
    python scripts/infer_launchlist.py > gpurun_out/ncu_infer.log 2>&1
echo "ncu exit $?"
