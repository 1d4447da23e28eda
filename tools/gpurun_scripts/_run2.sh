python -m pytest tests/test_gpu_ilu.py tests/test_gpu_multigrid.py tests/test_gpu_coarse.py -q --tb=short -x 2>&1 | grep -v Warn | tail -25 > gpurun_out/r2n_ilu_tests.log
cat gpurun_out/r2n_ilu_tests.log
timeout 300 python tools/mg_profile.py 1408 8 > gpurun_out/r2n_mg_profile_packed.log 2>&1
tail -14 gpurun_out/r2n_mg_profile_packed.log
