#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_wnfused_pair_n python tools/wn_single.py 1,25600,16 > gpurun_out/r2_ncu_n.log 2>&1
tail -3 gpurun_out/r2_ncu_n.log
