#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_t9_gemm.log
tail -12 gpurun_out/r02_t9_gemm.log
if grep -q "failed\|error\|Error" gpurun_out/r02_t9_gemm.log; then echo "GEMM TESTS FAILED - skipping bench"; exit 0; fi
timeout 600 python -m pytest tests/test_clip_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -s -k "not test_a and not test_c and not test_e" 2>&1 | grep -v Warning > gpurun_out/r02_t9_clip.log
VIT_B=95 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:(gemm_f16)' --launch-skip 100 -c 40 --csv --log-file gpurun_out/r02_t9_launch.csv \
    python tools/profile_step.py vit 0 > gpurun_out/ncu9.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline --skip-eager --skip-cpu --skip-voxel > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err
grep -n "(b)\|(d)\|passed\|failed\|FAILED" gpurun_out/r02_t9_clip.log | tail -12
cut -c1-300 gpurun_out/r02_bench_f.json
tail -3 gpurun_out/r02_bench_f.err
