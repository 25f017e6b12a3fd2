#!/bin/bash
# Source-level (SASS) sampling of two GEMM launches of a full pass: resblock2_1_conv1 (stride 2) and resblock2_1_conv2
set -u
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --utts 32 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_shift' -s 28 -c 2 -f -o /tmp/src $B > $OUT/src.log 2>&1
ncu -i /tmp/src.ncu-rep --page source --csv --print-source sass > $OUT/src_sass.csv 2>> $OUT/src.log
ncu -i /tmp/src.ncu-rep --page raw --csv > $OUT/src_raw.csv 2>> $OUT/src.log
ls -la $OUT/src*; tail -3 $OUT/src.log
