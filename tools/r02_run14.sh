#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/debug_halo_pair.py > gpurun_out/r02_dbg_halo.log 2>&1
tail -30 gpurun_out/r02_dbg_halo.log
