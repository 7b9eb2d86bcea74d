#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "halo" 2>&1 | tail -30 > gpurun_out/r02_t13_halo.log
tail -8 gpurun_out/r02_t13_halo.log
if grep -q "failed\|rror" gpurun_out/r02_t13_halo.log; then echo "HALO TESTS FAILED"; exit 0; fi
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_baseline_configs_gpu.py tests/test_properties_gpu.py -m gpu -q -s -k "not test_b and not test_d and not relevancy" 2>&1 | grep -v Warning > gpurun_out/r02_t13_unet.log
grep -n "(a)\|(c)\|(e)\|passed\|failed\|FAILED" gpurun_out/r02_t13_unet.log | tail -12
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv3d_halo' -c 12 --csv --log-file gpurun_out/r02_t13_launch.csv \
    python tools/profile_step.py unet 0 > gpurun_out/ncu13.log 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 --images 1 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-ours > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_g.json')); v=d['voxel']; print('voxel', v['value'], v['ms_per_step']); [print(k) for k in v['roofline']['kernels'][:4]]
PY
