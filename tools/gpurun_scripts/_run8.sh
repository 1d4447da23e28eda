python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -12 > gpurun_out/r2k_gputests.log; cat gpurun_out/r2k_gputests.log
python bench.py > gpurun_out/r2k_bench_1gpu.json 2> gpurun_out/r2k_bench_1gpu.err; tail -c 600 gpurun_out/r2k_bench_1gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2k_bench_2gpu_weak.json 2> gpurun_out/r2k_bench_2gpu_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/r2k_bench_2gpu_strong.json 2> gpurun_out/r2k_bench_2gpu_strong.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err
