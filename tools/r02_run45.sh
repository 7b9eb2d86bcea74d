#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py tests/test_clip_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_bwd_cls -c 12 --csv --log-file gpurun_out/r02_cls_launches.csv python bench.py --steps 1 --warmup 1 --images 1 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel --skip-ours > /dev/null 2>&1
grep attn_bwd_cls gpurun_out/r02_cls_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -12 | tr '\n' ' '; echo
timeout 900 python bench.py --steps 2 --warmup 3 --images 2 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel --skip-ours > gpurun_out/r02_bench_cls.json 2> gpurun_out/r02_bench_cls.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_cls.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac'])
PY
