#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/l_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/l_pytest_gpu.log
timeout 600 python scripts/r02/mode_times.py ${1:-mixed16} train > gpurun_out/k_modes_train.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_train.log | grep -v -i Warn
