#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multigrid.py -q -x --tb=short 2>&1 | grep -v "Warning\|torch.sparse\|^$" | tail -8 > gpurun_out/r2x_mg_tests.log
cat gpurun_out/r2x_mg_tests.log
python tools/mg_profile.py 2>&1 | grep -v "Warning\|torch.sparse" > gpurun_out/r2x_mg_profile.log
cat gpurun_out/r2x_mg_profile.log
python bench.py --no-cpu --steps 5 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2x_bench.json").read().strip().splitlines()[-1]); n=d["newton"]
print(n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"])
PY
