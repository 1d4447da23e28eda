#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2v_bench_2gpu_weak.json 2> gpurun_out/r2v_bench_2gpu_weak.err
tail -3 gpurun_out/r2v_bench_2gpu_weak.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2v_bench_2gpu_weak.json").read().strip().splitlines()[-1]); print(d["value"], d["spmv"]["ms"], d["e2e"]); print(json.dumps(d["newton"], indent=1))
PY
