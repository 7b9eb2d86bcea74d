#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_properties_gpu.py tests/test_pipeline.py -m gpu -q -s -k "unet or folded or halo or pipeline or ovssc" 2>&1 | grep -v Warning | grep "folded\|passed\|failed\|Error\|rror" | tail -8
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k "test_a_ or test_c_" 2>&1 | grep "(a)\|(c)\|passed\|failed\|rror" | tail
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_unet_traffic.csv python tools/profile_step.py unet 0 > gpurun_out/ncu37.log 2>&1
python tools/summarize_traffic.py gpurun_out/r02_unet_traffic.csv > gpurun_out/r02_unet_traffic_summary.txt; head -14 gpurun_out/r02_unet_traffic_summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 --images 1 --skip-pipeline --skip-eager --skip-cpu --skip-ours --skip-train > gpurun_out/r02_bench_p.json 2> gpurun_out/r02_bench_p.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_p.json')); v=d['voxel']; print('voxel', v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'], v['roofline'].get('traffic'), v['roofline'].get('frac')); [print(k) for k in v['roofline']['kernels'][:7]]
PY
tail -3 gpurun_out/r02_bench_p.err
