#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_properties_gpu.py tests/test_unet_bwd_gpu.py tests/test_train_gpu.py -m gpu -q 2>&1 | grep -v Warning | tail -3
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k "test_a_ or test_c_ or test_e_" 2>&1 | grep "(a)\|(c)\|(e)\|passed\|failed\|rror" | tail
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv3d_igemm' -c 80 --csv --log-file gpurun_out/r02_t40_launch.csv python tools/profile_step.py unet 0 > gpurun_out/ncu40.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_t40_launch.csv')) if len(r)>14 and r[0].isdigit()]
d={}
for r in rows: d.setdefault(int(r[0]),{'name':r[4].split('(')[0]})[r[12]]=float(r[14])
for k in sorted(d):
    if True: print(k, d[k]['name'][-30:], '%.1f us'%(d[k]['gpu__time_duration.sum']/1e3), '%.0f%%'%d[k].get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',0))
print('igemm total %.2f ms'%(sum(v['gpu__time_duration.sum'] for v in d.values())/1e6))
PY
timeout 900 python bench.py --steps 1 --warmup 1 --images 1 --skip-ours --skip-eager --skip-cpu --skip-pipeline --skip-train > gpurun_out/r02_bench_s.json 2> gpurun_out/r02_bench_s.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_s.json')); v=d['voxel']; print('rel', d['value'], 'voxel', v['value'], v['ms_per_step'], 'e2e', v['e2e']['value']); [print(k) for k in v['roofline']['kernels'][:6]]
PY
