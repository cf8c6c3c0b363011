#!/bin/bash
# round 2, call bc: ring depths of the fused EdgeConv forward (W stages x B slots): 3x3 (production), 3x4, 2x4, 2x5
mkdir -p gpurun_out
L=graphnet_b200/csrc
cp $L/libgraphnet_b200.so /tmp/orig.so
for v in orig w2b3 w2b2 w3b2; do
  if [ $v = orig ]; then cp /tmp/orig.so $L/libgraphnet_b200.so; else cp $L/libvariant_$v.so $L/libgraphnet_b200.so; fi
  timeout 600 python -m pytest tests/test_gpu_bf16.py -q -x -k "fused" > gpurun_out/bc_pytest_$v.log 2>&1; echo "$v pytest exit $?"
  timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/bc_mode_train_$v.log 2>&1; grep -v Warn gpurun_out/bc_mode_train_$v.log | grep "==\|agg_fused"
done
cp /tmp/orig.so $L/libgraphnet_b200.so
