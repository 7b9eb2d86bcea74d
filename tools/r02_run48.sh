#!/bin/bash
# steady-state DRAM traffic of ONE ResidualUNet3D forward (128^3 x 32 ch, batch 4): profiled window = the third forward
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
STEADY=1 SEMABS_UNET_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_unet_traffic_steady.csv python tools/profile_step.py unet 1 > gpurun_out/ncu48.log 2>&1
python tools/summarize_traffic.py gpurun_out/r02_unet_traffic_steady.csv > gpurun_out/r02_unet_traffic_steady_summary.txt; head -24 gpurun_out/r02_unet_traffic_steady_summary.txt
