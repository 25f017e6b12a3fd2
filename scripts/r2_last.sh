#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2l_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2l_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
SAN_TIMEOUT=420 bash scripts/sanitize.sh
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f frac %.3f stft %.3f istft %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["kernels"]["stft"]["frac_hbm"], d["kernels"]["istft"]["frac_hbm"], d["clocks"]))
PY
