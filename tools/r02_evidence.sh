#!/bin/bash
# Round-2 evidence run on a B200 box (gpurun): GPU test-suite, ncu --set full summaries (made on the box) of every kernel family
# with >= 2 % of the relevancy step and of the halo / igemm convolutions, ncu launch list of the bench command, DRAM traffic.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -4 > gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_gpu_tests.log
FILT='sm__ops_path\|sm__mem_tensor_cycles_active\.\(max\|min\|sum\)\|pipe_tensor_cycles_active\.\(max\|min\|sum\)\| 0 \| 0$\|TriageCompute\|device__attribute'
cap() { # name regex skip count top script args...
  local name=$1 rx=$2 skip=$3 cnt=$4 top=$5; shift 5
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu_ev_$name.log 2>&1
  { echo "# ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c $cnt $*"; python tools/ncu_summary.py /tmp/$name.ncu-rep 2>&1 | grep -v "$FILT"; python tools/ncu_hot.py /tmp/$name.ncu-rep $top 2>&1; } > gpurun_out/r02_ncu_full_$name.txt
  echo "$name: $(grep -c '^== ' gpurun_out/r02_ncu_full_$name.txt) sections"
}
export VIT_B=95
cap attn_bwd3 'attn_bwd_(row|col)_tc3|attn_bwd_tail2|attn_delta' 8 4 14 python tools/profile_step.py vit 0
cap attn_fwd 'attn_fwd_tc' 14 2 14 python tools/profile_step.py vit 0
cap gemm_pair_bench 'gemm_f16_tn_pair' 60 8 8 python tools/profile_step.py vit 0
cap ln_bwd 'layernorm_bwd' 20 2 10 python tools/profile_step.py vit 0
cap halo_pair 'conv3d_halo_pair' 3 2 12 python tools/profile_step.py unet 0
cap igemm 'conv3d_igemm' 20 4 10 python tools/profile_step.py unet 0
# launch list of the bench command (one image of the relevancy workload + the voxel section)
ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/r02_launch_bench.csv \
    python bench.py --steps 1 --warmup 1 --images 1 --skip-cpu --skip-train --skip-pipeline --skip-eager --skip-ours > gpurun_out/ncu_ev_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launch_bench.csv > gpurun_out/r02_launch_bench_summary.txt 2>&1; head -30 gpurun_out/r02_launch_bench_summary.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:gemm_f16_tn --launch-skip 100 -c 100 --csv --log-file gpurun_out/r02_gemm_traffic.csv python tools/profile_step.py vit 1 > gpurun_out/ncu_ev_gt.log 2>&1
python tools/summarize_traffic.py gpurun_out/r02_gemm_traffic.csv > gpurun_out/r02_gemm_traffic_summary.txt 2>&1; head -5 gpurun_out/r02_gemm_traffic_summary.txt
