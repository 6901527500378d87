#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stylemelgan.py tests/test_gpu_disc_loss.py -q -m gpu -s 2>&1 | grep -E "passed|failed|FAILED|STYLE-GRAD|Error|assert" | cut -c1-400
