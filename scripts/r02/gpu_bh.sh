#!/bin/bash
# round 2, call bh: programmatic dependent launch on the final build: which tests fail, what the step gains
mkdir -p gpurun_out
GNB_PDL=1 timeout 900 python -m pytest tests -q -m gpu > gpurun_out/bh_pytest_pdl.log 2>&1; echo "pytest exit $?"; grep -E "^FAILED|passed|failed" gpurun_out/bh_pytest_pdl.log | head -20
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bh_mode_train.log 2>&1; grep -v Warn gpurun_out/bh_mode_train.log | head -3
GNB_PDL=1 timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bh_mode_train_pdl.log 2>&1; grep -v Warn gpurun_out/bh_mode_train_pdl.log | head -3
