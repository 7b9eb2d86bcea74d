#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py tests/test_clip_gpu.py -m gpu -q -s 2>&1 | grep -v Warning | grep "err\|passed\|failed\|FAILED\|jitter" | tail -14
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k "test_b_ or test_d_" 2>&1 | grep "(b)\|(d)\|passed\|failed"
timeout 900 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel --skip-ours > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_j.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks']); r=d['roofline']; print(r['frac'], r['whole_path_frac']); [print(k) for k in r['kernels'][:5]]
PY
