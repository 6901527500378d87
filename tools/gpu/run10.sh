#!/bin/bash
# round 2, run 10: tile-pair kernel with TMA store / reduce epilogue
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_layer_packed or weight_update" > gpurun_out/r2_tests_k0.log 2>&1; tail -3 gpurun_out/r2_tests_i0.log
for v in 0 8 4 12 160; do PWGB_WN_VARIANT=$v timeout 120 python tools/wn_single.py 1,25600,16 2>&1 | grep -E "variant|FUSED"; done > gpurun_out/r2_wn_variants_k.txt
PWGB_WN_VARIANT=0 timeout 120 python tools/wn_single.py 512,25600,16 64,25600,16 1,25600,64 1,25600,1 2>&1 | grep -E "FUSED" >> gpurun_out/r2_wn_variants_k.txt
cat gpurun_out/r2_wn_variants_k.txt
timeout 120 python tools/wn_trace.py 1,25600,16 0 > gpurun_out/r2_wn_trace_k.txt 2>&1
grep -A14 "== gate\|== epi-skip\|== epi-x\|== conv0" gpurun_out/r2_wn_trace_k.txt | cut -c1-80
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -q -m gpu -k "c3 or pwg or decoder" > gpurun_out/r2_tests_k1.log 2>&1; tail -3 gpurun_out/r2_tests_i1.log
