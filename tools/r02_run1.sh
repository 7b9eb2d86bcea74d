#!/bin/bash
# round-2 GPU call 1: new BASELINE-config parity tests + ADVICE regression tests, then ncu --set full of the kernels
# that own the non-GEMM 45 % of the relevancy step (bench shapes: 95 tiles x 16 labels, ViT-L/14)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_baseline_configs_gpu.py tests/test_train_gpu.py -m gpu -q -s -x 2>&1 | tail -60 > gpurun_out/r02_tests1.log
RX='regex:(attn_fwd_tc|attn_bwd_row_tc|attn_bwd_col_tc|attn_bwd_tail|layernorm_bwd_warp|gemm_f16_tn_kernel)'
VIT_B=95 timeout 900 ncu --set full --clock-control none --import-source on -k "$RX" --launch-skip 150 -c 30 -f -o gpurun_out/r02_vit_bwd \
    python tools/profile_step.py vit 0 > gpurun_out/ncu_bwd.log 2>&1
VIT_B=95 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:(attn_fwd_tc|gemm_f16_tn_kernel|layernorm_fwd)' --launch-skip 40 -c 14 -f -o gpurun_out/r02_vit_fwd \
    python tools/profile_step.py vit 0 > gpurun_out/ncu_fwd.log 2>&1
cat gpurun_out/r02_tests1.log
ls -la gpurun_out/*.ncu-rep
