#!/bin/bash
mkdir -p gpurun_out
RUNS=30 timeout 600 python tools/debug_bwd_determinism.py 2>&1 | grep -v Warning | tail -80
