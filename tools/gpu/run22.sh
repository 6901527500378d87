#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r2_tests_r.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r2_tests_r.log | cut -c1-300
timeout 300 python tools/pwg_forward_bench.py 1 16 64 > gpurun_out/r2_pwg_forward_r.json 2> gpurun_out/r2_pwg_forward_r.err
grep -E '"ms"|samples_per_s|classes' gpurun_out/r2_pwg_forward_r.json | cut -c1-250
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_r.json 2> gpurun_out/r2_bench_r.err
tail -4 gpurun_out/r2_bench_r.err
