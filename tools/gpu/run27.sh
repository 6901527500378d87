#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2_bench_2gpu_f.json 2> gpurun_out/r2_bench_2gpu_f.err
tail -c 400 gpurun_out/r2_bench_2gpu_f.json; tail -4 gpurun_out/r2_bench_2gpu_f.err
