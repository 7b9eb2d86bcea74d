#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py -m gpu -q -x -s -k "attn_bwd" 2>&1 | grep -v Warning | grep "gen 3\|passed\|failed\|Error\|assert" | tail -12
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -5
export VIT_B=95
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:attn_' --launch-skip 60 -c 12 --csv --log-file gpurun_out/r02_t36_attn.csv python tools/profile_step.py vit 0 > gpurun_out/ncu36.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_t36_attn.csv')) if len(r)>14 and r[0].isdigit()]
for r in rows: print(r[4].split('(')[0][:40], r[14])
PY
timeout 900 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-ours > gpurun_out/r02_bench_o.json 2> gpurun_out/r02_bench_o.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_o.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks']); r=d['roofline']; print(r['frac'], r['whole_path_frac']); [print(k) for k in r['kernels'][:6]]
v=d['voxel']; print('voxel', v['value'], v['ms_per_step'], v['e2e']['value']); [print(k) for k in v['roofline']['kernels'][:6]]
PY
