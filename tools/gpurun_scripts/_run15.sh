#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multigrid.py -q -x --tb=short -s 2>&1 | tail -25 > gpurun_out/r2t_mg_tests.log
cat gpurun_out/r2t_mg_tests.log
python bench.py --no-cpu --steps 5 > gpurun_out/r2t_bench_aux.json 2> gpurun_out/r2t_bench_aux.err
tail -5 gpurun_out/r2t_bench_aux.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2t_bench_aux.json").read().strip().splitlines()[-1]); print(json.dumps(d["newton"], indent=1))
PY
