#!/bin/bash
cd $GRAFT_REPO_ROOT
ECH_KSP_PROFILE=1 python bench.py --no-cpu --steps 5 > gpurun_out/r2x_bench_prof.json 2> gpurun_out/r2x_bench_prof.err
grep "KSP profile" gpurun_out/r2x_bench_prof.err
python bench.py --no-cpu --steps 5 > gpurun_out/r2x_bench2.json 2> gpurun_out/r2x_bench2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2x_bench2.json").read().strip().splitlines()[-1]); n=d["newton"]
print(n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"])
PY
nproc; grep -m1 "model name" /proc/cpuinfo
