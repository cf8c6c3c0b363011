#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r02/profile_masked.py > gpurun_out/p_masked_plain.log 2>&1; echo "plain exit $?"; cat gpurun_out/p_masked_plain.log | grep -v Warn
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad_build" --launch-skip 3 --launch-count 1 \
   -o gpurun_out/p_wgrad_build -f python scripts/r02/profile_masked.py > gpurun_out/p_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/p_ncu.log
