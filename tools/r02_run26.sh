#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py -m gpu -q -x -s -k "attn_bwd" 2>&1 | grep -v Warning | grep "gen 3\|passed\|failed\|Error\|assert" | tail -15
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k "test_b_ or test_d_" 2>&1 | grep "(b)\|(d)\|passed\|failed"
VIT_B=95 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k 'regex:attn_' --launch-skip 60 -c 8 --csv --log-file gpurun_out/r02_t33_attn.csv \
    python tools/profile_step.py vit 0 > gpurun_out/ncu33.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_t33_attn.csv')) if len(r)>14 and r[0].isdigit()]
d={}
for r in rows:
    d.setdefault((r[0], r[4].split('(')[0]), {})[r[12].split('.')[0][-14:]]=r[14]
for k,v in d.items(): print(k[1][:40], v)
PY
timeout 900 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel --skip-ours > gpurun_out/r02_bench_n.json 2> gpurun_out/r02_bench_n.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks']); r=d['roofline']; print(r['frac'], r['whole_path_frac']); [print(k) for k in r['kernels'][:6]]
PY
