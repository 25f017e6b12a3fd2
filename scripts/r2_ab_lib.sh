#!/bin/bash
# A/B of two library builds inside one box: layer table of nhans_b200/libnhans_b200.so (new) and of $1 (old), interleaved
set -u
mkdir -p gpurun_out
OLD=${1:-nhans_b200/libnhans_b200_prev.so}
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/abl_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/abl_tests.log
for i in 1 2; do
  NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/abl_new_$i.txt 2>&1; grep -E "audio-s/s|all GEMM" gpurun_out/abl_new_$i.txt
  NHANS_B200_LIB=$PWD/$OLD NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/abl_old_$i.txt 2>&1; grep -E "audio-s/s|all GEMM" gpurun_out/abl_old_$i.txt
done
echo "layer  new(ms TFLOP/s)  old(ms TFLOP/s)"
paste <(grep TFLOP gpurun_out/abl_new_2.txt | awk '{print $1, $5, $(NF-1)}') <(grep TFLOP gpurun_out/abl_old_2.txt | awk '{print $5, $(NF-1)}')
grep -A20 "wait cycles" gpurun_out/abl_new_2.txt | head -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench new', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
NHANS_B200_LIB=$PWD/$OLD timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench old', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
