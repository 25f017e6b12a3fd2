#!/bin/bash
# quick GPU check after a kernel change: network / end-to-end tests, per-layer table with wait statistics, one bench line
set -u
mkdir -p gpurun_out
T=${1:-chk}
timeout 1500 python -m pytest tests/test_gpu_net.py tests/test_gpu_e2e.py tests/test_gpu_dsp.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/${T}_layers.txt 2>&1; cat gpurun_out/${T}_layers.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - "$T" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/%s_bench.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("bench value %.1f e2e %.1f ms/step %.2f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]),
          "stft %.3f istft %.3f" % (d["kernels"]["stft"]["frac_hbm"], d["kernels"]["istft"]["frac_hbm"]), d["clocks"])
except Exception as ex:
    print("bench parse failed", ex)
PY
