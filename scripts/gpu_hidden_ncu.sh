#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_hidden_fwd_node --launch-skip 3 --launch-count 1 \
   -o gpurun_out/r_hidden_fwd -f python scripts/hidden_probe.py > gpurun_out/ncu_hidden.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_hidden.log
