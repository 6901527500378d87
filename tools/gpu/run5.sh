set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 31; do PWGB_WN_VARIANT=$v timeout 120 python tools/wn_single.py 1,25600,16 512,25600,16 1,25600,64 2>&1 | grep -E "variant|FUSED"; done > gpurun_out/r2_wn_variants_e.txt
cat gpurun_out/r2_wn_variants_e.txt
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r2_tests_e.log 2>&1
grep -v "Warn\|warn" gpurun_out/r2_tests_e.log | tail -25
timeout 300 python tools/pwg_forward_bench.py 16 > gpurun_out/r2_pwg_forward_e.json 2> gpurun_out/r2_pwg_forward_e.err
grep -E '"ms"|samples_per_s|frac' gpurun_out/r2_pwg_forward_e.json
timeout 600 python tools/train_profile.py > gpurun_out/r2_train_profile_e.txt 2>&1
grep -v "Warn\|warn" gpurun_out/r2_train_profile_e.txt | head -30
timeout 900 python bench.py --steps 10 --warmup 3 --no-eager > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
tail -c 400 gpurun_out/r2_bench_e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_e.json'))
for k in ['value','ms_per_step','parity','train','train_pwg']:
    print(k, json.dumps(d.get(k))[:400])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_wnfused_d1_e python tools/wn_single.py 1,25600,16 > /dev/null 2>&1
ls gpurun_out | tail -3
