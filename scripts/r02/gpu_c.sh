#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config0.py -q -s > gpurun_out/c_pytest_config0.log 2>&1; echo "config0 exit $?"
grep -E "batch of|prometheus50|passed|failed" gpurun_out/c_pytest_config0.log | head -40
timeout 900 python -m pytest tests/test_gpu_users.py -q -s > gpurun_out/c_pytest_users.log 2>&1; echo "users exit $?"
grep -E "out |passed|failed|Error" gpurun_out/c_pytest_users.log | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1; echo "smoke exit $?"; grep -E "smoke\[|Error|assert" gpurun_out/c_smoke.log | head
timeout 2400 python -m pytest tests -q -m gpu -s --deselect tests/test_gpu_config0.py --deselect tests/test_gpu_users.py > gpurun_out/c_pytest_gpu.log 2>&1; echo "pytest rest exit $?"; tail -6 gpurun_out/c_pytest_gpu.log
grep -E "train_step|golden |default config|config #4|FAILED" gpurun_out/c_pytest_gpu.log | head -60
