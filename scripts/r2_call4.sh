#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2c4_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "parity\]|passed|failed|Error" gpurun_out/r2c4_tests.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c4_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c4_smoke.log
for c in cfg2 cfg1 cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c4_bench_$c.json 2> gpurun_out/r2c4_bench_$c.err; echo "bench $c rc=$?"
done
python - <<'PY'
import json
for c in ("cfg2","cfg1","cfg3","cfg4","cfg5"):
    try:
        d = json.loads(open("gpurun_out/r2c4_bench_%s.json" % c).read().strip().splitlines()[-1])
        print(c, "value %.1f e2e %.1f ms/step %.2f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]),
              "stft %.2f istft %.2f" % (d["kernels"]["stft"]["frac_hbm"], d["kernels"]["istft"]["frac_hbm"]), d["clocks"]["sm_mhz"], d.get("latency_ms"))
    except Exception as ex:
        print(c, "parse failed", ex)
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2c4_ref.json 2> gpurun_out/r2c4_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2c4_ref.json
