#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python tools/train_profile.py torchprof > gpurun_out/r2_train_profile_hifigan_torch.txt 2>&1
grep -v "Warn\|warn" gpurun_out/r2_train_profile_hifigan_torch.txt | cut -c1-92,170-290 | head -60
