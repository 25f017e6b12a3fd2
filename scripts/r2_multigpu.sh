#!/bin/bash
# 8-GPU box: the driver-style torchrun bench at N = 2 and 8 (cfg5 shard slices) and BASELINE config 5 in full through
# runtime.MultiGpu (8192 x 10 s, dynamic chunk dealing), plus its single-GPU reference point (1024 x 10 s on one GPU).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 2 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2mg_bench_n$N.json 2> gpurun_out/r2mg_bench_n$N.err; echo "bench N=$N rc=$?"
done
timeout 300 python bench.py --gpus 1 --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2mg_bench_n1.json 2> gpurun_out/r2mg_bench_n1.err; echo "bench N=1 cfg5 rc=$?"
python - <<'PY'
import json
for n in (1, 2, 8):
    try:
        d = json.loads(open("gpurun_out/r2mg_bench_n%d.json" % n).read().strip().splitlines()[-1])
        print("N=%d value %.1f e2e %.1f ms/step %.1f per-rank" % (n, d["value"], d["e2e"]["value"], d["ms_per_step"]), [round(x, 1) for x in d["per_rank_ms_per_step"]], d["config"]["name"])
    except Exception as ex:
        print(n, "parse failed", ex)
PY
timeout 900 python scripts/cfg5_multigpu.py --gpus 8 --utts 8192 --seconds 10 --chunk 64 --repeat 2 > gpurun_out/r2mg_cfg5_8gpu.json 2> gpurun_out/r2mg_cfg5_8gpu.err; echo "multigpu 8 rc=$?"; cut -c1-600 gpurun_out/r2mg_cfg5_8gpu.json
timeout 600 python scripts/cfg5_multigpu.py --gpus 1 --utts 1024 --seconds 10 --chunk 64 --repeat 1 > gpurun_out/r2mg_cfg5_1gpu.json 2> gpurun_out/r2mg_cfg5_1gpu.err; echo "multigpu 1 rc=$?"; cut -c1-400 gpurun_out/r2mg_cfg5_1gpu.json
