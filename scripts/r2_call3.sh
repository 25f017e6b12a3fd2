#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dsp.py tests/test_gpu_net.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/r2c3_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r2c3_tests.log
NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2c3_layers_walk.txt 2>&1; echo "walk rc=$?"
head -6 gpurun_out/r2c3_layers_walk.txt; grep -A4 "wait cycles" gpurun_out/r2c3_layers_walk.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c3_bench.json").read().strip().splitlines()[-1])
    print("bench value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"])
    print("kernels", json.dumps(d["kernels"]))
except Exception as ex:
    print("bench parse failed", ex)
PY
B="python bench.py --steps 1 --warmup 1 --utts 32 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv64_walk -s 6 -c 3 -f -o /tmp/r2c3_walk $B > gpurun_out/r2c3_ncu_walk.log 2>&1
ncu -i /tmp/r2c3_walk.ncu-rep --page raw --csv > gpurun_out/r2c3_walk_raw.csv 2>> gpurun_out/r2c3_ncu_walk.log
timeout 600 ncu --set full --clock-control none -k regex:'stft_kernel|istft_kernel' -c 6 -f -o /tmp/r2c3_dsp python bench.py --steps 1 --warmup 1 --utts 256 --no-cpu-baseline > gpurun_out/r2c3_ncu_dsp.log 2>&1
ncu -i /tmp/r2c3_dsp.ncu-rep --page raw --csv > gpurun_out/r2c3_dsp_raw.csv 2>> gpurun_out/r2c3_ncu_dsp.log
ls -la gpurun_out | grep r2c3
