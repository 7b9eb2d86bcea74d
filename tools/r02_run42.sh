#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_properties_gpu.py -m gpu -q -k "unet or conv_transpose" 2>&1 | grep -v Warning | tail -4
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k "test_a_" 2>&1 | grep "(a)\|passed\|failed\|rror" | tail -4
SEMABS_UNET_GRAPH=1 timeout 900 python bench.py --steps 1 --warmup 1 --images 1 --skip-ours --skip-eager --skip-cpu --skip-pipeline --skip-train > gpurun_out/r02_bench_z.json 2> gpurun_out/r02_bench_z.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_z.json')); v=d['voxel']; print('voxel', v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'], 'launches', v['gpu_launches']); [print(k) for k in v['roofline']['kernels'][:7]]
PY
