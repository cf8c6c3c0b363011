#!/bin/bash
# round 2, call z: weight-gradient builder with two k_in tiles per CTA: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_train_step.py -q -x > gpurun_out/z_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/z_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/z_mode_train.log 2>&1; grep -v Warn gpurun_out/z_mode_train.log | head -14
