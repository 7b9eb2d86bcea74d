#!/bin/bash
mkdir -p gpurun_out
export VIT_B=95
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py -m gpu -q -x -k "attn_bwd" 2>&1 | grep -v Warning | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:attn_bwd_(row|col)_tc3' --launch-skip 6 -c 2 -f -o /tmp/attn3 python tools/profile_step.py vit 0 > gpurun_out/ncu32.log 2>&1
python tools/ncu_summary.py /tmp/attn3.ncu-rep 2>&1 | grep -v "sm__ops_path\| 0 \| 0$" > gpurun_out/r02_ncu_full_attn_bwd3_v7.txt
python tools/ncu_hot.py /tmp/attn3.ncu-rep 600 >> gpurun_out/r02_ncu_full_attn_bwd3_v7.txt 2>&1
grep "time_duration\|inst_executed.sum \|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" gpurun_out/r02_ncu_full_attn_bwd3_v7.txt
