#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_properties_gpu.py tests/test_pipeline.py -m gpu -q -k "unet or conv_transpose or pipeline or ovssc" 2>&1 | grep -v Warning | tail -4
for G in 1 0; do
SEMABS_UNET_GRAPH=$G timeout 900 python bench.py --steps 1 --warmup 1 --images 1 --skip-ours --skip-eager --skip-cpu --skip-pipeline --skip-train > gpurun_out/r02_bench_t$G.json 2> gpurun_out/r02_bench_t$G.err
python - $G <<'PY'
import json,sys
d=json.load(open('gpurun_out/r02_bench_t%s.json'%sys.argv[1])); v=d['voxel']; print('graph', sys.argv[1], 'voxel', v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'], 'launches', v['gpu_launches'])
PY
done
