#!/bin/bash
# final round-2 evidence run on one B200: full GPU tests, bench lines for every config, reference arm, layer table,
# ncu captures, compute-sanitizer
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2f_tests.log 2>&1; echo "gpu tests rc=$?"; grep -E "parity\]|passed|failed" gpurun_out/r2f_tests.log | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_cfg2.json 2> gpurun_out/r2f_bench_cfg2.err; echo "bench cfg2 rc=$?"
for c in cfg1 cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_$c.json 2> gpurun_out/r2f_bench_$c.err; echo "bench $c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for c in ("cfg2","cfg1","cfg3","cfg4","cfg5"):
    try:
        d = json.loads(open("gpurun_out/r2f_bench_%s.json" % c).read().strip().splitlines()[-1])
        print(c, "value %.1f e2e %.1f ms/step %.2f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]),
              "stft %.2f istft %.2f" % (d["kernels"]["stft"]["frac_hbm"], d["kernels"]["istft"]["frac_hbm"]), d["clocks"]["sm_mhz"], d.get("latency_ms"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("dedup_value"))
    except Exception as ex:
        print(c, "parse failed", ex)
try:
    d = json.loads(open("gpurun_out/r2f_ref.json").read().strip().splitlines()[-1]); print("ref", d["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("dedup_value"))
except Exception as ex:
    print("ref parse failed", ex)
PY
NHANS_DEBUG_STATS=1 timeout 300 python scripts/layer_profile.py 64 4 > gpurun_out/r2f_layers.txt 2>&1; head -3 gpurun_out/r2f_layers.txt
bash scripts/profile_capture.sh r2f
SAN_TIMEOUT=420 bash scripts/sanitize.sh
