#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_bwd_tail2 --launch-skip 4 -c 10 --csv --log-file gpurun_out/r02_tail2_launches.csv python tools/profile_step.py vit 0 > /dev/null 2>&1
grep attn_bwd_tail2 gpurun_out/r02_tail2_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -12 | tr '\n' ' '; echo
