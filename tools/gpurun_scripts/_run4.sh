( time python bench.py > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err ) 2> gpurun_out/r2g_bench_1gpu.time
tail -3 gpurun_out/r2g_bench_1gpu.err; cat gpurun_out/r2g_bench_1gpu.time
( time python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err ) 2> gpurun_out/r2g_bench_ref.time
tail -3 gpurun_out/r2g_bench_ref.err; cat gpurun_out/r2g_bench_ref.time
