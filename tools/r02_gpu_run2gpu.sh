python -m pytest tests/test_gpu_ddp.py -x -q > gpurun_out/g15_ddp.log 2>&1; tail -3 gpurun_out/g15_ddp.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary > gpurun_out/g15_bench_2gpu.json 2> gpurun_out/g15_bench_2gpu.err; echo rc=$?; tail -1 gpurun_out/g15_bench_2gpu.json | cut -c1-300
