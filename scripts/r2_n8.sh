#!/bin/bash
# 8-GPU record of the last kernel revision: driver-style torchrun bench at N = 8 and BASELINE config 5 in full through runtime.MultiGpu
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err; echo "bench N=8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n8_bench.json").read().strip().splitlines()[-1])
print("N=8 value %.1f e2e %.1f ms/step %.1f per-rank" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), [round(x, 1) for x in d["per_rank_ms_per_step"]], d["config"]["name"], d["clocks"])
PY
timeout 600 python scripts/cfg5_multigpu.py --gpus 8 --utts 8192 --seconds 10 --chunk 32 --repeat 1 > gpurun_out/r2n8_cfg5.json 2> gpurun_out/r2n8_cfg5.err; echo "multigpu 8 rc=$?"; cut -c1-1500 gpurun_out/r2n8_cfg5.json
