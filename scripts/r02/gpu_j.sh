#!/bin/bash
# round 2, call j: bf16-plane kernels bring-up
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x ${1:+-k "$1"} > gpurun_out/j_pytest_bf16.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/j_pytest_bf16.log
