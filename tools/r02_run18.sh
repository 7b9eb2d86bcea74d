#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_unet_bwd_gpu.py tests/test_properties_gpu.py tests/test_pipeline.py -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/r02_t18_unet.log
grep -n "passed\|failed\|FAILED" gpurun_out/r02_t18_unet.log | tail -8
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k "test_a_ or test_c_ or test_e_" 2>&1 | grep -v Warning > gpurun_out/r02_t18_base.log
grep -n "^(a)\|^(c)\|^(e)\|(a)\|(c)\|(e)\|passed\|failed\|FAILED" gpurun_out/r02_t18_base.log | tail -12
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv3d_' -c 80 --csv --log-file gpurun_out/r02_t18_launch.csv \
    python tools/profile_step.py unet 0 > gpurun_out/ncu18.log 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 --images 1 --skip-pipeline --skip-eager --skip-cpu --skip-ours --skip-amp > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_h.json')); v=d['voxel']; print('voxel', v['value'], v['ms_per_step']); [print(k) for k in v['roofline']['kernels'][:4]]
t=d['train']; print('train ms', t['ms_per_step'], t['loss_trajectory'])
PY
tail -3 gpurun_out/r02_bench_h.err
