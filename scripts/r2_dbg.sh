#!/bin/bash
# timing experiments with the GEMM kernel's debug switches (garbage results): which resource bounds each layer?
set -u
mkdir -p gpurun_out
for m in 0 64 128 192 256 448; do
  NHANS_DESC_MODE=$m timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/dbg_$m.txt 2>&1
  echo "== mode $m"; grep -E "audio-s/s|all GEMM" gpurun_out/dbg_$m.txt
done
paste <(grep TFLOP gpurun_out/dbg_0.txt | awk '{print $1, $5}') <(grep TFLOP gpurun_out/dbg_64.txt | awk '{print $5}') <(grep TFLOP gpurun_out/dbg_128.txt | awk '{print $5}') <(grep TFLOP gpurun_out/dbg_192.txt | awk '{print $5}') <(grep TFLOP gpurun_out/dbg_256.txt | awk '{print $5}') <(grep TFLOP gpurun_out/dbg_448.txt | awk '{print $5}')
