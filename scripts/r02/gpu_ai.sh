#!/bin/bash
# round 2, call ai: max aggregation in the GEMM epilogue + arg-routed backward (EdgeConvTito route)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tito_max.py -q -x -s > gpurun_out/ai_pytest_max.log 2>&1; echo "pytest max exit $?"; tail -5 gpurun_out/ai_pytest_max.log
timeout 900 python -m pytest tests/test_gpu_users.py tests/test_gpu_tc.py tests/test_gpu_ops.py -q -x -s > gpurun_out/ai_pytest.log 2>&1; echo "pytest exit $?"; grep -E "tito|passed|failed" gpurun_out/ai_pytest.log | tail -8
