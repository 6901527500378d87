#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -q -m gpu -x 2>&1 | grep -E "passed|failed"
PWGB_WN_VARIANT=0 timeout 120 python tools/wn_single.py 1,25600,16 1,25600,64 2>&1 | grep FUSED
