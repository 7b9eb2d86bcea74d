#!/bin/bash
# round-2 GPU call 3: fixed tests + a SMALL ncu rep (2 kernels, source-level stall sampling) of the attention backward
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s -k test_e 2>&1 | grep -v Warning | tail -30 > gpurun_out/r02_tests_e.log
timeout 1500 python -m pytest tests/test_train_gpu.py tests/test_train_boundary_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02_tests_train.log
VIT_B=95 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:attn_bwd_(row|col)_tc' --launch-skip 6 -c 2 -f -o gpurun_out/r02_attn_bwd \
    python tools/profile_step.py vit 0 > gpurun_out/ncu_attn.log 2>&1
VIT_B=95 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_f16_tn_kernel' --launch-skip 165 -c 4 -f -o gpurun_out/r02_gemm_bwd \
    python tools/profile_step.py vit 0 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -12 gpurun_out/r02_tests_e.log
tail -8 gpurun_out/r02_tests_train.log
