#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/debug_halo_pair.py > gpurun_out/r02_dbg_halo.log 2>&1
grep -n "time-out\|waits\|match\|Error" gpurun_out/r02_dbg_halo.log | head -60
grep -c "ok" gpurun_out/r02_dbg_halo.log
