#!/bin/bash
# round 2, run 17: 2-GPU bench sanity (own arm + reference arm) and the re-bounded PWG gradient test
cd /root/repo; mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_ddp.py -q -m gpu -k "pwg_train_step or ddp" 2>&1 | grep -E "passed|failed|BAD" | cut -c1-400
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -c 1500 gpurun_out/r2_bench_2gpu.json; tail -4 gpurun_out/r2_bench_2gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/r2_bench_ref_2gpu.json 2> gpurun_out/r2_bench_ref_2gpu.err
tail -c 800 gpurun_out/r2_bench_ref_2gpu.json; tail -4 gpurun_out/r2_bench_ref_2gpu.err
