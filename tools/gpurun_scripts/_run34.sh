#!/bin/bash
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2z_bench_8gpu_weak.json 2> gpurun_out/r2z_bench_8gpu_weak.err
python - <<'PY'
import json
for f in ["gpurun_out/r2z_bench_8gpu_weak.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); n=d["newton"]
        print(f, d["value"], d["ms_per_step"], d["spmv"]["ms"], d["e2e"]["value"], n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"])
    except Exception as e:
        print(f, "ERR", e)
PY
grep -v "Warning\|sparse_csr" gpurun_out/r2z_bench_8gpu_weak.err | tail -3
