#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/r02/mode_times.py ${1:-tf32x3,bf16x3,bf16} train > gpurun_out/k_modes_train.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_train.log | grep -v -i Warn
timeout 600 python scripts/r02/mode_times.py ${2:-tf32,bf16} infer > gpurun_out/k_modes_infer.log 2>&1; echo "modes exit $?"; cat gpurun_out/k_modes_infer.log | grep -v -i Warn
