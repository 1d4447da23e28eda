python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_ilu.py -q --tb=short -x 2>&1 | grep -v Warn | tail -12
nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dmma_probe.cu -o /tmp/dmma_probe && /tmp/dmma_probe > gpurun_out/r2p_dmma_probe.txt 2>&1; cat gpurun_out/r2p_dmma_probe.txt
timeout 300 python tools/mg_profile.py 1408 8 2>&1 | grep -v Warn | tail -13 > gpurun_out/r2p_mg_profile.log; head -3 gpurun_out/r2p_mg_profile.log
python bench.py > gpurun_out/r2p_bench_1gpu.json 2> gpurun_out/r2p_bench_1gpu.err; tail -c 200 gpurun_out/r2p_bench_1gpu.json
