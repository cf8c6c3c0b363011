#!/bin/bash
# kernel role probe only: bash scripts/gpu_probe.sh <mode>
mkdir -p gpurun_out
timeout 300 python scripts/linear_probe.py $1 > gpurun_out/probe_$1.log 2>&1; echo "probe exit $?"; cat gpurun_out/probe_$1.log
