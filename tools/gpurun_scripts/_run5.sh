python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2h_bench_2gpu_weak.json 2> gpurun_out/r2h_bench_2gpu_weak.err
tail -2 gpurun_out/r2h_bench_2gpu_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/r2h_bench_2gpu_strong.json 2> gpurun_out/r2h_bench_2gpu_strong.err
tail -2 gpurun_out/r2h_bench_2gpu_strong.err
