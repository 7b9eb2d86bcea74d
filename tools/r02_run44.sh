#!/bin/bash
mkdir -p gpurun_out
for TB in 95 143 285; do
SEMABS_TILE_BATCH=$TB timeout 900 python bench.py --steps 2 --warmup 3 --images 2 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel --skip-ours > gpurun_out/r02_bench_tb$TB.json 2> gpurun_out/r02_bench_tb$TB.err
python - $TB <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r02_bench_tb%s.json'%sys.argv[1])); print('tile batch', sys.argv[1], 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac'])
except Exception as e:
    print('tile batch', sys.argv[1], 'failed', e); print(open('gpurun_out/r02_bench_tb%s.err'%sys.argv[1]).read()[-600:])
PY
done
