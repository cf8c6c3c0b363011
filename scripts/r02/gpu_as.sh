#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r02/deg9_stats.py > gpurun_out/as_deg9.log 2>&1; grep -v Warn gpurun_out/as_deg9.log | tail -12
