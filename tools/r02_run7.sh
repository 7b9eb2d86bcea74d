#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py -m gpu -q -k "attn_bwd" 2>&1 | tail -15 > gpurun_out/r02_t7_kern.log
timeout 600 python -m pytest tests/test_clip_gpu.py -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/r02_t7_clip.log
VIT_B=95 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:(attn_|layernorm_bwd|gemm_f16)' --launch-skip 160 -c 70 --csv --log-file gpurun_out/r02_t7_launch.csv \
    python tools/profile_step.py vit 0 > gpurun_out/ncu7.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
tail -6 gpurun_out/r02_t7_kern.log
grep -n "color jitter\|ours\|passed\|failed\|FAILED" gpurun_out/r02_t7_clip.log | tail -25
cut -c1-300 gpurun_out/r02_bench_d.json
