#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3; do
  GNB_WGRAD_SWAP=$v timeout 200 python -m pytest tests/test_gpu_tc.py -q -k wgrad > gpurun_out/wgrad_v$v.log 2>&1
  echo "variant $v: $(tail -1 gpurun_out/wgrad_v$v.log)"
done
