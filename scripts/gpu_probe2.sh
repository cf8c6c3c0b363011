#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/hidden_probe.py > gpurun_out/probe_hidden.log 2>&1; echo "probe exit $?"; cat gpurun_out/probe_hidden.log
