python -m pytest tests/test_gpu_ilu.py tests/test_gpu_multigrid.py tests/test_gpu_coarse.py -q --tb=short -x 2>&1 | tail -25 > gpurun_out/r2e_ilu_tests.log
cat gpurun_out/r2e_ilu_tests.log
python tools/mg_profile.py 1408 8 > gpurun_out/r2e_mg_profile_planned.log 2>&1
tail -14 gpurun_out/r2e_mg_profile_planned.log
