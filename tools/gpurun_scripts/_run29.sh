#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -v "Warning\|torch.sparse\|^$" | tail -6 > gpurun_out/r2z_gpu_suite.log
cat gpurun_out/r2z_gpu_suite.log
python bench.py > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench_1gpu.err
tail -c 600 gpurun_out/r2z_bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
cat gpurun_out/r2z_bench_reference.json | head -c 800
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
