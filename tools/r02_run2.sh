#!/bin/bash
# round-2 GPU call 2: remaining new tests, first bench run with the new legs, ncu --set full summaries (text only: the
# .ncu-rep files exceed gpurun's 64 MiB return limit, so they are summarised on the box)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/r02_tests_baseline.log
timeout 1500 python -m pytest tests/test_train_gpu.py tests/test_train_boundary_gpu.py -m gpu -q -s -x 2>&1 | tail -40 > gpurun_out/r02_tests_train.log
timeout 1500 python bench.py --steps 2 --warmup 3 --skip-train --skip-pipeline > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
RX='regex:(attn_fwd_tc|attn_bwd_row_tc|attn_bwd_col_tc|attn_bwd_tail|layernorm_bwd_warp|gemm_f16_tn_kernel)'
VIT_B=95 timeout 900 ncu --set full --clock-control none -k "$RX" --launch-skip 150 -c 12 -f -o /tmp/r02_vit_bwd \
    python tools/profile_step.py vit 0 > gpurun_out/ncu_bwd.log 2>&1
python tools/ncu_summary.py /tmp/r02_vit_bwd.ncu-rep > gpurun_out/r02_ncu_full_vit_bwd.txt 2>&1
VIT_B=95 timeout 900 ncu --set full --clock-control none -k 'regex:(attn_fwd_tc|gemm_f16_tn_kernel|layernorm_fwd)' --launch-skip 40 -c 8 -f -o /tmp/r02_vit_fwd \
    python tools/profile_step.py vit 0 > gpurun_out/ncu_fwd.log 2>&1
python tools/ncu_summary.py /tmp/r02_vit_fwd.ncu-rep > gpurun_out/r02_ncu_full_vit_fwd.txt 2>&1
ls -la /tmp/*.ncu-rep
tail -5 gpurun_out/r02_tests_baseline.log gpurun_out/r02_tests_train.log
cut -c1-1500 gpurun_out/r02_bench_a.json
tail -5 gpurun_out/r02_bench_a.err
