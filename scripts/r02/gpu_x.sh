#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r02/gaps.py mixed16 > gpurun_out/x_gaps.log 2>&1; grep -v Warn gpurun_out/x_gaps.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/x_mode_train.log 2>&1; grep -v Warn gpurun_out/x_mode_train.log | head -12
timeout 300 python scripts/knn_probe.py 512 2>&1 | tail -3
