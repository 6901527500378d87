#!/bin/bash
# round 2, run 7: timeline of the fused WaveNet kernel + the re-bounded gradient tests
cd /root/repo; mkdir -p gpurun_out
set -x
for v in 0 8 4 12; do timeout 120 python tools/wn_trace.py 1,25600,16 $v > gpurun_out/r2_wn_trace_v$v.txt 2>&1; done
head -120 gpurun_out/r2_wn_trace_v0.txt
timeout 900 python -m pytest tests/test_gpu_backward.py -q -m gpu -x -s -k "hifigan_train_step or causal_hifigan or pwg_train_step" > gpurun_out/r2_tests_g.log 2>&1
grep -E "passed|failed|BAD|loose|worst" gpurun_out/r2_tests_g.log | cut -c1-1500
