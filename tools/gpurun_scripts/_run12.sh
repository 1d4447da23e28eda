free -g | head -2
timeout 600 python tools/bench_3d.py 2 2>&1 | grep "3D Bortels" | tee gpurun_out/r2q_bench_cfg4_3d.txt
timeout 600 python tools/bench_cg.py 1600x800 2>&1 | grep "CG Gupta" | tee gpurun_out/r2q_bench_cfg3_cg.txt
timeout 900 python tools/bench_cg.py 3200x1600 2>&1 | grep "CG Gupta\|Error\|error" | tee -a gpurun_out/r2q_bench_cfg3_cg.txt
python bench.py --degree 2 --no-newton --no-cpu --steps 10 > gpurun_out/r2q_bench_cfg2_Q2.json 2>/dev/null
python bench.py --degree 3 --no-newton --no-cpu --steps 10 > gpurun_out/r2q_bench_cfg2_Q3.json 2>/dev/null
python -c "
import json
for q in (2,3):
    d=json.loads([l for l in open('gpurun_out/r2q_bench_cfg2_Q%d.json'%q) if l.startswith('{')][-1])
    print('Q%d'%q, d['ms_per_step'], d['roofline']['frac'], d['spmv']['ms'], d['spmv']['frac'])
"
