#!/bin/bash
cd $GRAFT_REPO_ROOT
for t in 8x8 16x4 32x2; do
  python bench.py --no-cpu --steps 5 --tile $t > gpurun_out/r2u_tile_$t.json 2> gpurun_out/r2u_tile_$t.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2u_tile_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); n=d["newton"]
        print(f, n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"], d["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
