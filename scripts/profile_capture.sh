#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list + one full capture of a 2048-window pass of the tensor-core layers
# (conv64_walk_kernel x 3, gemm_shift_kernel x 14) + the DSP kernels.  The .ncu-rep files stay in /tmp (they exceed
# the gpurun_out size limit); only CSV exports come back.
# Usage: gpurun --timeout 1500 -- 'bash scripts/profile_capture.sh TAG'   ->  gpurun_out/TAG_*.csv
set -u
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --utts 32 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
# 14 tower launches (Silent.wav embedding + the batch's contexts) and the first mask-network pass are skipped
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_shift|conv64_walk' -s 31 -c 17 -f -o /tmp/${TAG}_tc $B > $OUT/${TAG}_tc.log 2>&1
ncu -i /tmp/${TAG}_tc.ncu-rep --page raw --csv > $OUT/${TAG}_tc_raw.csv 2>> $OUT/${TAG}_tc.log
if [ -z "${SKIP_DSP_NCU:-}" ]; then
timeout 600 ncu --set full --clock-control none -k regex:'stft_kernel|istft_kernel' -c 8 -f -o /tmp/${TAG}_dsp python bench.py --steps 1 --warmup 1 --utts 256 --no-cpu-baseline > $OUT/${TAG}_dsp.log 2>&1
ncu -i /tmp/${TAG}_dsp.ncu-rep --page raw --csv > $OUT/${TAG}_dsp_raw.csv 2>> $OUT/${TAG}_dsp.log
fi
ls -la $OUT | grep ${TAG}
