#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py tests/test_clip_gpu.py tests/test_baseline_configs_gpu.py -m gpu -x -q -s 2>&1 | grep -i "passed\|failed\|error\|assert\|(b) " | tail -12
timeout 900 python bench.py --steps 2 --warmup 3 --images 2 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel --skip-ours > gpurun_out/r02_bench_fwdtail.json 2> gpurun_out/r02_bench_fwdtail.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_fwdtail.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac'], d['roofline'].get('whole_path_frac'))
for k in d['roofline'].get('kernels', [])[:5]: print(k)
PY
