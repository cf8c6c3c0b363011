#!/bin/bash
# round 2, call be: where the idle stream time of a step sits; every kernel on the maximum shared-memory carveout
mkdir -p gpurun_out
timeout 300 python scripts/r02/gaps.py mixed16 > gpurun_out/be_gaps.log 2>&1; grep -v "Warn\|warn" gpurun_out/be_gaps.log | head -40
GNB_CARVEOUT=1 timeout 300 python scripts/r02/gaps.py mixed16 > gpurun_out/be_gaps_carve.log 2>&1; grep -v "Warn\|warn" gpurun_out/be_gaps_carve.log | head -40
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/be_mode_train.log 2>&1; grep -v Warn gpurun_out/be_mode_train.log | head -4
GNB_CARVEOUT=1 timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/be_mode_train_carve.log 2>&1; grep -v Warn gpurun_out/be_mode_train_carve.log | head -4
