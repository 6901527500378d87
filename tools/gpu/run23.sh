#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_layer_packed and (515 or 700)" > gpurun_out/r2_memcheck_wn.log 2>&1; echo "memcheck wn rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/r2_memcheck_wn.log | head -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -q -m gpu -x -k "pwg_train_step" > gpurun_out/r2_memcheck_bwd.log 2>&1; echo "memcheck bwd rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/r2_memcheck_bwd.log | head -8
