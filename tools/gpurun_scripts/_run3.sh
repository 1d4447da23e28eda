for r in refine_never refine_ifneeded refine_always; do
ECH_REFINE=$r python tools/dist_mg_check.py 256 > gpurun_out/r2f_mg_1gpu_256_$r.log 2>&1; tail -1 gpurun_out/r2f_mg_1gpu_256_$r.log
done
python tools/dist_mg_check.py 1408 > gpurun_out/r2f_mg_1gpu_1408.log 2>&1; tail -1 gpurun_out/r2f_mg_1gpu_1408.log
python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_coarse.py -q --tb=short 2>&1 | tail -4
