#!/bin/bash
# 3D per-equation neighbour accumulation: parity + timing; tile-shape probe for the Newton step
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -k "3d or 3D or extruded or bortels" --tb=short 2>&1 | tail -8 > gpurun_out/r2s_3d_tests.log
python tools/bench_3d.py 2 > gpurun_out/r2s_bench_3d.log 2>&1
for t in 32x2 64x1 16x4 128x1; do
  python bench.py --no-cpu --steps 5 --tile $t > gpurun_out/r2s_tile_$t.json 2> gpurun_out/r2s_tile_$t.err
done
tail -3 gpurun_out/r2s_3d_tests.log gpurun_out/r2s_bench_3d.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s_tile_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); n=d["newton"]
        print(f, n["newton_step_s"], n["ksp_its_per_step"], n["timers_s_max_over_ranks"]["PCSetUp"], d["value"], d["spmv"])
    except Exception as e:
        print(f, "ERR", e)
PY
