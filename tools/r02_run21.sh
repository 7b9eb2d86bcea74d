#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -s -k "folded or halo" 2>&1 | grep -v Warning | grep "folded\|passed\|failed\|Error\|rror" | tail -12
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py tests/test_properties_gpu.py -m gpu -q -s -k "test_a_ or unet" 2>&1 | grep "(a)\|passed\|failed\|rror" | tail
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_unet_traffic.csv python tools/profile_step.py unet 0 > gpurun_out/ncu21.log 2>&1
python tools/summarize_traffic.py gpurun_out/r02_unet_traffic.csv > gpurun_out/r02_unet_traffic_summary.txt; head -12 gpurun_out/r02_unet_traffic_summary.txt
timeout 900 python bench.py --steps 1 --warmup 1 --images 1 --skip-pipeline --skip-eager --skip-cpu --skip-ours --skip-train > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_k.json')); v=d['voxel']; print('voxel', v['value'], v['ms_per_step']); [print(k) for k in v['roofline']['kernels'][:5]]
PY
tail -3 gpurun_out/r02_bench_k.err
