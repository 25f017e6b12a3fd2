#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dsp.py tests/test_gpu_e2e.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/dsp_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/dsp_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/dsp_bench.json 2> gpurun_out/dsp_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/dsp_bench.json").read().strip().splitlines()[-1])
k = d["kernels"]
print("value %.1f frac %.3f clocks %s" % (d["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
print("stft %.3f ms %.3f of HBM | istft %.3f ms %.3f of HBM" % (k["stft"]["ms"] / k["stft"]["launches"], k["stft"]["frac_hbm"], k["istft"]["ms"] / k["istft"]["launches"], k["istft"]["frac_hbm"]))
PY
timeout 600 ncu --set full --clock-control none -k regex:'stft_kernel|istft_kernel' -c 6 -f -o /tmp/dspn python bench.py --steps 1 --warmup 1 --utts 256 --no-cpu-baseline > gpurun_out/dsp_ncu.log 2>&1
ncu -i /tmp/dspn.ncu-rep --page raw --csv > gpurun_out/dsp_ncu_raw.csv 2>> gpurun_out/dsp_ncu.log
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/dsp_ncu_raw.csv")))
h = rows[0]
for r in rows[2:5]:
    print(r[h.index("Kernel Name")][:40], r[h.index("gpu__time_duration.sum")], "us  inst", r[h.index("smsp__inst_executed.sum")], " issue", r[h.index("smsp__issue_active.avg.pct_of_peak_sustained_active")][:5], " dram%", r[h.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")][:5])
PY
