#!/bin/bash
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2zz_bench_2gpu_weak.json 2> gpurun_out/r2zz_bench_2gpu_weak.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2zz_bench_2gpu_weak.json").read().strip().splitlines()[-1]); n=d["newton"]
print(d["value"], d["ms_per_step"], d["spmv"]["ms"], d["e2e"]["value"], n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"])
PY
