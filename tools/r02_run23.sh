#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py -m gpu -q -x -k "attn_bwd" 2>&1 | grep -v Warning | tail -15
