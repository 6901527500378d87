#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for v in 0 512 1024 1536 32 1568; do PWGB_WN_VARIANT=$v timeout 120 python tools/wn_single.py 1,25600,16 2>&1 | grep -E "variant|FUSED"; done
