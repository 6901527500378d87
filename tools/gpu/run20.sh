#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python tools/host_profile.py > gpurun_out/r2_host_profile.txt 2>&1
grep -v "Warn\|warn" gpurun_out/r2_host_profile.txt | cut -c1-150 | head -110
