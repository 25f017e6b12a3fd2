#!/bin/bash
# round 2, GPU call 1: validate the row-walk kernel, then compare the three executions of the 64-channel stage
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -x -q -m gpu > gpurun_out/r2c1_net.log 2>&1; echo "net tests rc=$?"
tail -15 gpurun_out/r2c1_net.log
NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2c1_layers_walk.txt 2>&1; echo "walk rc=$?"
NHANS_STAGE1=pair NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2c1_layers_pair.txt 2>&1; echo "pair rc=$?"
NHANS_NO_WALK=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2c1_layers_plain.txt 2>&1; echo "plain rc=$?"
head -12 gpurun_out/r2c1_layers_walk.txt; head -6 gpurun_out/r2c1_layers_pair.txt; head -6 gpurun_out/r2c1_layers_plain.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c1_all.log 2>&1; echo "all gpu tests rc=$?"
tail -5 gpurun_out/r2c1_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c1_bench.json").read().strip().splitlines()[-1])
    print("bench value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"])
except Exception as ex:
    print("bench parse failed", ex)
PY
