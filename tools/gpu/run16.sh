#!/bin/bash
# round 2, run 16: full validation of the pair kernel + bench + profiles
cd /root/repo; mkdir -p gpurun_out
set -x
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r2_tests_p.log 2>&1
grep -v "Warn\|warn" gpurun_out/r2_tests_p.log | grep -E "passed|failed|FAILED|BAD|Error" | cut -c1-600
PWGB_WN_VARIANT=0 timeout 200 python tools/wn_single.py 1,25600,16 2,25600,16 16,25600,16 64,25600,16 128,25600,16 512,25600,16 1,25600,64 1,25600,4 1,25600,1 2>&1 | grep -E "FUSED|2-launch" > gpurun_out/r2_wn_layers_p.txt
cat gpurun_out/r2_wn_layers_p.txt
timeout 300 python tools/pwg_forward_bench.py 1 16 64 > gpurun_out/r2_pwg_forward_p.json 2> gpurun_out/r2_pwg_forward_p.err
grep -E '"ms"|samples_per_s|frac' gpurun_out/r2_pwg_forward_p.json
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_p.json 2> gpurun_out/r2_bench_p.err
tail -c 3000 gpurun_out/r2_bench_p.json; tail -5 gpurun_out/r2_bench_p.err
timeout 120 python tools/wn_trace.py 1,25600,16 0 > gpurun_out/r2_wn_trace_p.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_wnfused_final python tools/wn_single.py 1,25600,16 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench_p.csv python bench.py --steps 2 --warmup 1 --no-eager > gpurun_out/r2_bench_under_ncu_p.log 2>&1
ls -la gpurun_out | tail -5
