CMD="python bench.py --steps 3 --warmup 3 --no-newton --no-cpu"
$CMD > gpurun_out/r2m_plain.json 2> gpurun_out/r2m_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches.csv $CMD > gpurun_out/r2m_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobian_coop2 -s 4 -c 1 -o gpurun_out/prof_r2m $CMD > gpurun_out/r2m_ncu.log 2>&1
tail -3 gpurun_out/r2m_ncu.log
python bench.py > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.err; tail -c 400 gpurun_out/r2m_bench_1gpu.json
