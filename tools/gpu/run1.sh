set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
# new fused kernel first, under its own timeout (a protocol bug traps, it cannot hang the box)
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "wavenet_fused" ) > gpurun_out/r2_tests_fused.log 2>&1
tail -15 gpurun_out/r2_tests_fused.log
timeout 300 python tools/wn_single.py 1,25600,16 16,25600,16 128,25600,16 512,25600,16 > gpurun_out/r2_wn_single_a.txt 2>&1
cat gpurun_out/r2_wn_single_a.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests_a.log 2>&1
tail -8 gpurun_out/r2_tests_a.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 600 gpurun_out/r2_bench_a.err
timeout 300 python tools/tc_single.py 128,11,1,25600,16 128,3,1,25600,16 64,3,1,51200,16 >> gpurun_out/r2_wn_single_a.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_tc_c128k11 python tools/tc_single.py 128,11,1,25600,16 > /dev/null 2>&1
WN_FUSED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -s 6 -c 2 -o gpurun_out/r2_ncu_wn2launch_d1 python tools/wn_single.py 1,25600,16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_wnfused_d1 python tools/wn_single.py 1,25600,16 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_a.csv python bench.py --steps 1 --warmup 3 --no-train --no-eager --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
