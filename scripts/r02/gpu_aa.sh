#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r02/fused_roles.py > gpurun_out/aa_fused_roles.log 2>&1; grep -v Warn gpurun_out/aa_fused_roles.log
