"""Time the assembly kernels of the cfg2 workload for several build variants (run on the GPU box)."""
import os
import sys
import subprocess
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = sys.argv[1:] or ["", "-DECH_JMINB=6", "-DECH_JMINB=8", "-DECH_COOP_WARPS=4", "-DECH_COOP_WARPS=4 -DECH_JMINB=3",
                            "-DECH_COOP_WARPS=1 -DECH_JMINB=12"]
for v in variants:
    env = dict(os.environ, ECH_MODULE_FLAGS=v)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu"],
                       env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        print("%-45s J %.3f ms  F %.3f ms  step %.3f ms" % (v or "(default)", j["jacobian_ms"], j["residual_ms"], j["ms_per_step"]), flush=True)
    except Exception as e:
        print(v, "FAILED", r.stdout[-500:], r.stderr[-1500:], flush=True)
