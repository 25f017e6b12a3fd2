#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2c5_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "parity\]|passed|failed|Error" gpurun_out/r2c5_tests.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c5_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c5_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c5_bench_cfg2.json 2> gpurun_out/r2c5_bench_cfg2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c5_bench_cfg2.json").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f ms/step %.2f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]),
          "stft %.3f istft %.3f" % (d["kernels"]["stft"]["frac_hbm"], d["kernels"]["istft"]["frac_hbm"]), d["clocks"])
    print(d["kernels"])
except Exception as ex:
    print("parse failed", ex)
PY
timeout 600 ncu --set full --clock-control none -k regex:'stft_kernel|istft_kernel' -c 6 -f -o /tmp/r2c5_dsp python bench.py --steps 1 --warmup 1 --utts 256 --no-cpu-baseline > gpurun_out/r2c5_ncu_dsp.log 2>&1
ncu -i /tmp/r2c5_dsp.ncu-rep --page raw --csv > gpurun_out/r2c5_dsp_raw.csv 2>> gpurun_out/r2c5_ncu_dsp.log
