#!/bin/bash
mkdir -p gpurun_out
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/r02_launch_bench.csv \
    python bench.py --steps 1 --warmup 1 --images 1 --skip-cpu --skip-train --skip-pipeline --skip-eager --skip-ours > gpurun_out/ncu_ev_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launch_bench.csv > gpurun_out/r02_launch_bench_summary.txt 2>&1; head -16 gpurun_out/r02_launch_bench_summary.txt
