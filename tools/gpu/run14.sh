#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 120 python tools/wn_trace.py 1,25600,16 0 > gpurun_out/r2_wn_trace_m.txt 2>&1
grep -A8 "== loader\|== conv0\|== conv1" gpurun_out/r2_wn_trace_m.txt | cut -c1-80
timeout 120 python tools/wn_trace.py 1,25600,16 12 > gpurun_out/r2_wn_trace_m12.txt 2>&1
grep -A8 "== loader\|== conv0" gpurun_out/r2_wn_trace_m12.txt | cut -c1-80
