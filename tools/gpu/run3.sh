set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strided_tc.py tests/test_gpu_optim.py -q -x -k "wavenet_fused or strided or optim or weight_updates or decoder or graphed" ) > gpurun_out/r2_tests_c.log 2>&1
tail -12 gpurun_out/r2_tests_c.log
for v in 0 1 2 3 4 8 12 16 19 31; do PWGB_WN_VARIANT=$v WN_ONLY_FUSED=1 timeout 120 python tools/wn_single.py 1,25600,16 2>&1 | grep -E "variant|FUSED"; done > gpurun_out/r2_wn_variants_c.txt
cat gpurun_out/r2_wn_variants_c.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_wnfused_d1_c python tools/wn_single.py 1,25600,16 > /dev/null 2>&1
ls -la gpurun_out | tail -5
