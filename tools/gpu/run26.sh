#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "other_channel_counts" 2>&1 | grep -E "passed|failed|FAILED|Error|assert" | cut -c1-400
