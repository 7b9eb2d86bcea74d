#!/bin/bash
# round-2 GPU call 5: second-generation attention backward (now actually selected), store tests, full logs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py tests/test_relevancy_store_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_t5_kern.log
timeout 1200 python -m pytest tests/test_clip_gpu.py tests/test_baseline_configs_gpu.py tests/test_train_gpu.py tests/test_train_boundary_gpu.py -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/r02_t5_e2e.log
VIT_B=95 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:(attn_|layernorm_bwd|gemm_f16)' --launch-skip 160 -c 70 --csv --log-file gpurun_out/r02_t5_launch.csv \
    python tools/profile_step.py vit 0 > gpurun_out/ncu5.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
tail -12 gpurun_out/r02_t5_kern.log
grep -n "(e)\|passed\|failed\|FAILED" gpurun_out/r02_t5_e2e.log | tail -12
cut -c1-400 gpurun_out/r02_bench_c.json
tail -3 gpurun_out/r02_bench_c.err
