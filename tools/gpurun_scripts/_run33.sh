#!/bin/bash
cd $GRAFT_REPO_ROOT
for eta in 0.05 0.005; do
ECH_GMRES_REFINE_ETA2=$eta timeout 600 python bench.py --no-cpu --steps 5 > gpurun_out/r2z_bench_eta_$eta.json 2> gpurun_out/r2z_bench_eta_$eta.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2z_bench_eta_$eta.json").read().strip().splitlines()[-1]); n=d["newton"]
print("$eta", n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"], n["max_nodal_error_vs_exact_fields"], n["ksp_reductions_last_solve"])
PY
done
