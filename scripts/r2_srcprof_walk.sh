#!/bin/bash
# SASS-level sampling of the row-walk kernel with the operand generator (resblock1_1_conv2, second pass of a 32 x 4 s step)
set -u
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --utts 32 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv64_walk' -s 3 -c 1 -f -o /tmp/walk $B > $OUT/walksrc.log 2>&1
ncu -i /tmp/walk.ncu-rep --page source --csv --print-source sass > $OUT/walksrc_sass.csv 2>> $OUT/walksrc.log
ls -la $OUT/walksrc*; head -3 $OUT/walksrc_sass.csv | cut -c1-200
