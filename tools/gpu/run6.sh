set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "wavenet_fused or pwg" ) > gpurun_out/r2_tests_f0.log 2>&1
tail -5 gpurun_out/r2_tests_f0.log
for v in 0 31; do PWGB_WN_VARIANT=$v timeout 120 python tools/wn_single.py 1,25600,16 512,25600,16 1,25600,64 2>&1 | grep -E "variant|FUSED"; done > gpurun_out/r2_wn_variants_f.txt
cat gpurun_out/r2_wn_variants_f.txt
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r2_tests_f.log 2>&1
grep -v "Warn\|warn" gpurun_out/r2_tests_f.log | tail -15
timeout 300 python tools/pwg_forward_bench.py 1 16 64 > gpurun_out/r2_pwg_forward_f.json 2> gpurun_out/r2_pwg_forward_f.err
grep -E '"ms"|samples_per_s|frac' gpurun_out/r2_pwg_forward_f.json
timeout 600 python tools/train_profile.py > gpurun_out/r2_train_profile_f.txt 2>&1
grep -v "Warn\|warn" gpurun_out/r2_train_profile_f.txt | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_wnfused_d1_f python tools/wn_single.py 1,25600,16 > /dev/null 2>&1
ls gpurun_out | tail -3
