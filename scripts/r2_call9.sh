#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_net.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2c9_tests.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r2c9_tests.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f e2e %.1f ms/step %.2f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]),
          "stft %.3f (%.3f ms) istft %.3f (%.3f ms)" % (d["kernels"]["stft"]["frac_hbm"], d["kernels"]["stft"]["ms"] / d["kernels"]["stft"]["launches"], d["kernels"]["istft"]["frac_hbm"], d["kernels"]["istft"]["ms"] / d["kernels"]["istft"]["launches"]),
          "direct %.2f ms/step" % (d["kernels"]["direct_conv"]["ms"] / d["steps"]), "last_dense %.2f" % d["layers"][-1]["ms_per_step"], d["clocks"]["sm_mhz"])
except Exception as ex:
    print(sys.argv[1], "parse failed", ex)
PY
}
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err; show gpurun_out/r2c9_bench.json
NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2c9_layers.txt 2>&1; cat gpurun_out/r2c9_layers.txt
