#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for v in 0 32 128 160 256 288; do PWGB_WN_VARIANT=$v timeout 120 python tools/wn_single.py 1,25600,16 2>&1 | grep -E "variant|FUSED"; done > gpurun_out/r2_wn_variants_j.txt
cat gpurun_out/r2_wn_variants_j.txt
