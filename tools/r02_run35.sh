#!/bin/bash
mkdir -p gpurun_out
export VIT_B=95
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py -m gpu -q -x -k "attn_bwd" 2>&1 | grep -v Warning | tail -2
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:attn_bwd_(row|col)' --launch-skip 8 -c 8 --csv --log-file gpurun_out/r02_t35_attn.csv python tools/profile_step.py vit 0 > gpurun_out/ncu35.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_t35_attn.csv')) if len(r)>14 and r[0].isdigit()]
d={}
for r in rows:
    d.setdefault((r[0], r[4].split('(')[0]), {})[r[12].split('.')[0][-14:]]=r[14]
for k,v in d.items(): print(k[1][:40], v)
PY
