set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strided_tc.py tests/test_gpu_optim.py tests/test_gpu_baseline_shapes.py -q -x -k "wavenet_fused or strided or optim or weight_updates or pwg" ) > gpurun_out/r2_tests_b.log 2>&1
tail -15 gpurun_out/r2_tests_b.log
timeout 300 python tools/wn_single.py 1,25600,16 512,25600,16 1,25600,64 > gpurun_out/r2_wn_single_b.txt 2>&1
cat gpurun_out/r2_wn_single_b.txt
timeout 300 python tools/pwg_forward_bench.py 1 16 64 > gpurun_out/r2_pwg_forward_b.json 2> gpurun_out/r2_pwg_forward_b.err
cat gpurun_out/r2_pwg_forward_b.json
timeout 600 python tools/train_profile.py > gpurun_out/r2_train_profile_b.txt 2>&1
head -60 gpurun_out/r2_train_profile_b.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_wnfused_d1_b python tools/wn_single.py 1,25600,16 > /dev/null 2>&1
ls -la gpurun_out
