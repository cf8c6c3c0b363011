#!/bin/bash
# round 2, call t: kNN (FIFO-insertion scan) parity + timing + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -q -x > gpurun_out/t_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/t_pytest.log
timeout 300 python scripts/knn_probe.py 512 > gpurun_out/t_knn_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/t_knn_probe.log
timeout 300 python scripts/knn_probe.py 1024 >> gpurun_out/t_knn_probe.log 2>&1; tail -4 gpurun_out/t_knn_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_table_split --launch-skip 3 --launch-count 1 \
   -o gpurun_out/t_knn_split -f python scripts/knn_probe.py 512 > gpurun_out/t_ncu_knn.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/t_ncu_knn.log
