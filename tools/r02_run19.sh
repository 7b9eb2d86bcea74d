#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_unet_bwd_gpu.py tests/test_properties_gpu.py -m gpu -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k "test_a_ or test_c_" 2>&1 | grep "^(a)\|^(c)\|^\.(\|passed\|failed"
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv3d_' -c 80 --csv --log-file gpurun_out/r02_t19_launch.csv \
    python tools/profile_step.py unet 0 > gpurun_out/ncu19.log 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 --images 1 --skip-pipeline --skip-eager --skip-cpu --skip-ours --skip-train > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_i.json')); v=d['voxel']; print('voxel', v['value'], v['ms_per_step']); [print(k) for k in v['roofline']['kernels'][:4]]
PY
