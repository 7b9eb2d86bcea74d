#!/bin/bash
mkdir -p gpurun_out
VIT_B=95 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_f16_tn_pair' --launch-skip 104 -c 2 -f -o gpurun_out/r02_gemm_pair \
    python tools/profile_step.py vit 0 > gpurun_out/ncu10.log 2>&1
ls -la gpurun_out/r02_gemm_pair.ncu-rep
timeout 600 python -m pytest tests/test_pipeline.py tests/test_relevancy_store_gpu.py -m gpu -q 2>&1 | tail -5
