#!/bin/bash
# round-2 GPU call 4: second-generation attention backward — kernel tests, engine parity, per-kernel times, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py -m gpu -q -x -k "attn_bwd" 2>&1 | tail -25 > gpurun_out/r02_t4_kern.log
timeout 1200 python -m pytest tests/test_clip_gpu.py tests/test_baseline_configs_gpu.py tests/test_train_gpu.py tests/test_train_boundary_gpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -60 > gpurun_out/r02_t4_e2e.log
VIT_B=95 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:(attn_|layernorm_bwd|gemm_f16)' --launch-skip 160 -c 60 --csv --log-file gpurun_out/r02_t4_launch.csv \
    python tools/profile_step.py vit 0 > gpurun_out/ncu4.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
cat gpurun_out/r02_t4_kern.log | tail -8
tail -15 gpurun_out/r02_t4_e2e.log
cut -c1-700 gpurun_out/r02_bench_b.json
tail -3 gpurun_out/r02_bench_b.err
