#!/bin/bash
# SASS-level executed-instruction counts of the DSP kernels (one launch each of stft<phasor> and istft<phasor>)
set -u
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --utts 256 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stft_kernel' -s 3 -c 3 -f -o /tmp/dsp $B > $OUT/dspsrc.log 2>&1
ncu -i /tmp/dsp.ncu-rep --page source --csv --print-source sass > $OUT/dspsrc_sass.csv 2>> $OUT/dspsrc.log
ncu -i /tmp/dsp.ncu-rep --page source --csv --print-source cuda > $OUT/dspsrc_cuda.csv 2>> $OUT/dspsrc.log
ls -la $OUT/dspsrc*; grep -c . $OUT/dspsrc_sass.csv
