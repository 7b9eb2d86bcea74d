#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py tests/test_properties_gpu.py tests/test_unet_bwd_gpu.py tests/test_train_gpu.py tests/test_pipeline.py tests/test_train_boundary_gpu.py -m gpu -q 2>&1 | grep -v Warning | tail -5
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k "test_a_ or test_c_ or test_e_" 2>&1 | grep "(a)\|(c)\|(e)\|passed\|failed\|rror" | tail -8
timeout 900 python bench.py --steps 1 --warmup 1 --images 1 --skip-ours --skip-eager --skip-cpu --skip-pipeline > gpurun_out/r02_bench_y.json 2> gpurun_out/r02_bench_y.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_y.json')); v=d['voxel']; print('voxel', v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'], 'launches', v['gpu_launches']); [print(k) for k in v['roofline']['kernels'][:7]]
t=d['train']; print('train ms', t['ms_per_step'], t.get('amp_like',{}).get('ms_per_step'))
PY
