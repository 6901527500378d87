#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_layer_packed" 2>&1 | tail -1
for v in 0 8 4 12; do PWGB_WN_VARIANT=$v timeout 120 python tools/wn_single.py 1,25600,16 2>&1 | grep -E "variant|FUSED"; done
timeout 120 python tools/wn_trace.py 1,25600,16 0 > gpurun_out/r2_wn_trace_l.txt 2>&1
grep -A10 "== gate\|== epi-skip\|== epi-x\|== conv0" gpurun_out/r2_wn_trace_l.txt | cut -c1-80
