#!/bin/bash
# what the driver runs at round end, on one B200: GPU tests, smoke, both bench arms with its step counts
set -u
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" = 0 ]; then
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2d_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2d_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
fi
S=$SECONDS; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2d_ref.json 2> gpurun_out/r2d_ref.err; echo "ref rc=$? wall $((SECONDS-S)) s"
S=$SECONDS; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$? wall $((SECONDS-S)) s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r2d_ref.json").read().strip().splitlines()[-1])
print("ours value %.1f e2e %.1f ms/step %.1f frac %.3f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print("traffic", d["roofline"]["traffic"], d["roofline"]["traffic_source"])
print("ref value %.3f (%s)" % (r["value"], r["cpu_baseline"]["sample"]), "ratio e2e %.0f" % (d["e2e"]["value"] / r["value"]))
PY
