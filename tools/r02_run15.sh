#!/bin/bash
mkdir -p gpurun_out
SEMABS_HALO_PAIRS=2 timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "halo and True" 2>&1 | tail -30 > gpurun_out/r02_t15_halo.log
grep -n "time-out\|passed\|failed" gpurun_out/r02_t15_halo.log | head -20
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/debug_halo_pair.py > gpurun_out/r02_dbg_halo.log 2>&1
grep -n "time-out\|ok\|match\|Error" gpurun_out/r02_dbg_halo.log | head -30
