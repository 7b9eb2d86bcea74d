#!/bin/bash
mkdir -p gpurun_out
VIT_B=95 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:attn_(bwd_row_tc2|bwd_col_tc2|bwd_tail2|delta)' --launch-skip 8 -c 4 -f -o gpurun_out/r02_attn_bwd2 \
    python tools/profile_step.py vit 0 > gpurun_out/ncu6.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_clip_gpu.py -m gpu -q -s -k "jitter" 2>&1 | grep -v Warning | tail -15
