#!/bin/bash
# round 2, run 19: final validation + bench of the round-2 build
cd /root/repo; mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r2_tests_t.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r2_tests_t.log | cut -c1-300
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_t.json 2> gpurun_out/r2_bench_t.err
tail -c 600 gpurun_out/r2_bench_t.json; tail -4 gpurun_out/r2_bench_t.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
