#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/r2c11_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2c11_tests.log
NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2c11_layers_na6.txt 2>&1; head -22 gpurun_out/r2c11_layers_na6.txt; grep -A8 "wait cycles" gpurun_out/r2c11_layers_na6.txt | tail -5
NHANS_NA=4 NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2c11_layers_na4.txt 2>&1; head -10 gpurun_out/r2c11_layers_na4.txt
