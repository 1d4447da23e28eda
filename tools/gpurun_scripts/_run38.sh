#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -v "Warning\|torch.sparse\|^$" | tail -6 > gpurun_out/r2zz_gpu_suite.log
cat gpurun_out/r2zz_gpu_suite.log
python bench.py > gpurun_out/r2zz_bench_1gpu.json 2> gpurun_out/r2zz_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2zz_bench_1gpu.json").read().strip().splitlines()[-1]); n=d["newton"]
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["spmv"]["ms"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"])
print(n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"])
PY
