#!/bin/bash
# round 2, run 8: which L2->SM stream matters (timing variants of the two-pipeline kernel)
cd /root/repo; mkdir -p gpurun_out
for v in 0 128 2 130 1 3 131 8 4 16; do PWGB_WN_VARIANT=$v timeout 120 python tools/wn_single.py 1,25600,16 2>&1 | grep -E "variant|FUSED"; done > gpurun_out/r2_wn_variants_g.txt
cat gpurun_out/r2_wn_variants_g.txt
timeout 120 python tools/wn_trace.py 1,25600,16 128 > gpurun_out/r2_wn_trace_v128.txt 2>&1
tail -30 gpurun_out/r2_wn_trace_v128.txt
