#!/bin/bash
# A/B of a debug switch inside one box: layer table with and without NHANS_DESC_MODE=$1, twice each, interleaved
set -u
mkdir -p gpurun_out
M=${1:-32}
for i in 1 2; do
  NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/ab_on_$i.txt 2>&1; grep -E "audio-s/s|all GEMM" gpurun_out/ab_on_$i.txt
  NHANS_DESC_MODE=$M NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/ab_off_$i.txt 2>&1; grep -E "audio-s/s|all GEMM" gpurun_out/ab_off_$i.txt
done
paste <(grep TFLOP gpurun_out/ab_on_2.txt | awk '{print $1, $5, $NF, $(NF-1)}') <(grep TFLOP gpurun_out/ab_off_2.txt | awk '{print $5, $(NF-1)}')
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/ab_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/ab_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
NHANS_DESC_MODE=$M timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench off', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
