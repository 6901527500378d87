#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_ddp.py tests/test_gpu_parity.py -q -m gpu 2>&1 | grep -E "passed|failed|FAILED" | cut -c1-300
timeout 900 python tools/train_profile.py pwg 64 torchprof > gpurun_out/r2_train_profile_pwg_torch7.txt 2>&1
grep -v "Warn\|warn" gpurun_out/r2_train_profile_pwg_torch7.txt | cut -c1-70,150-260 | head -32
timeout 600 python tools/host_profile.py 2>&1 | grep "step ms"
