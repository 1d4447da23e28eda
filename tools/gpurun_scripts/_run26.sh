#!/bin/bash
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2y_bench_2gpu_weak.json 2> gpurun_out/r2y_bench_2gpu_weak.err
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --scaling strong > gpurun_out/r2y_bench_2gpu_strong.json 2> gpurun_out/r2y_bench_2gpu_strong.err
python - <<'PY'
import json
for f in ["gpurun_out/r2y_bench_2gpu_weak.json", "gpurun_out/r2y_bench_2gpu_strong.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); n=d["newton"]
        print(f, d["value"], d["ms_per_step"], d["spmv"]["ms"], d["e2e"]["value"], n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2y_bench_2gpu_strong.err
