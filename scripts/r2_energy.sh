#!/bin/bash
set -u
mkdir -p gpurun_out
for m in 0 64 128 192 256 0; do
  NHANS_DESC_MODE=$m timeout 300 python scripts/energy_probe.py 64 4 24 2>&1 | tail -1 | tee -a gpurun_out/energy.txt
done
