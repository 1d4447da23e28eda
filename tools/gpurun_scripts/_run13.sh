timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2r_bench_8gpu_weak.json 2> gpurun_out/r2r_bench_8gpu_weak.err
echo "exit $?"; tail -c 300 gpurun_out/r2r_bench_8gpu_weak.json
