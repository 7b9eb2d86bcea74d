#!/bin/bash
# full GPU suite + ncu --set full evidence of the kernels that stay as they are this round (summaries made on the box)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -6 > gpurun_out/r02_t22_tests.log
cat gpurun_out/r02_t22_tests.log
cap() { # name regex skip count script args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu22_$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/r02_ncu_full_$name.txt 2>&1
  python tools/ncu_hot.py /tmp/$name.ncu-rep 12 >> gpurun_out/r02_ncu_full_$name.txt 2>&1
  grep -c "^==" gpurun_out/r02_ncu_full_$name.txt
}
export VIT_B=95
cap ln_bwd 'layernorm_bwd' 20 2 python tools/profile_step.py vit 0
cap gemm_pair_bench 'gemm_f16_tn_pair' 150 8 python tools/profile_step.py vit 0
cap attn_fwd 'attn_fwd_tc' 30 2 python tools/profile_step.py vit 0
cap halo_pair 'conv3d_halo_pair' 3 2 python tools/profile_step.py unet 0
ls -la /tmp/*.ncu-rep
