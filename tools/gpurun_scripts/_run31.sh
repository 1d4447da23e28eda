#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_ilu.py tests/test_gpu_multigrid.py -q -x --tb=short 2>&1 | grep -v "Warning\|torch.sparse\|^$" | tail -12
timeout 300 python tools/mg_profile.py 2>&1 | grep "V-cycle total\|whole set-up\|set-up parts: factor"
timeout 600 python bench.py --no-cpu --steps 5 > gpurun_out/r2z_bench_b.json 2> gpurun_out/r2z_bench_b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench_b.json").read().strip().splitlines()[-1]); n=d["newton"]
print(n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"], n["max_nodal_error_vs_exact_fields"])
PY
