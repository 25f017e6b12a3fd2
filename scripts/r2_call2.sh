#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -x -q -m gpu > gpurun_out/r2c2_net.log 2>&1; echo "net tests rc=$?"
tail -5 gpurun_out/r2c2_net.log
NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2c2_layers_walk.txt 2>&1; echo "walk rc=$?"
cat gpurun_out/r2c2_layers_walk.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c2_bench.json").read().strip().splitlines()[-1])
    print("bench value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"])
except Exception as ex:
    print("bench parse failed", ex)
PY
