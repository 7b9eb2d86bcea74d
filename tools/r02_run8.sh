#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_t8_kern.log
timeout 600 python -m pytest tests/test_clip_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -s -k "not test_a and not test_c and not test_e" 2>&1 | grep -v Warning > gpurun_out/r02_t8_clip.log
VIT_B=95 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:(attn_|layernorm_bwd|gemm_f16)' --launch-skip 160 -c 70 --csv --log-file gpurun_out/r02_t8_launch.csv \
    python tools/profile_step.py vit 0 > gpurun_out/ncu8.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err
tail -6 gpurun_out/r02_t8_kern.log
grep -n "(b)\|(d)\|passed\|failed\|FAILED" gpurun_out/r02_t8_clip.log | tail -12
cut -c1-300 gpurun_out/r02_bench_e.json
tail -3 gpurun_out/r02_bench_e.err
