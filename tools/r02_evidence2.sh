#!/bin/bash
# Round-2 evidence, second part (after the voxel changes): UNet traffic + launch list, ncu --set full of the new voxel kernels,
# final bench lines of both arms, micro-benchmarks, attention trace.  Every capture is bounded.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -3 > gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_gpu_tests.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
SEMABS_UNET_GRAPH=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_unet_traffic.csv python tools/profile_step.py unet 1 > gpurun_out/ncu_ev2_traffic.log 2>&1
python tools/summarize_traffic.py gpurun_out/r02_unet_traffic.csv > gpurun_out/r02_unet_traffic_summary.txt; head -16 gpurun_out/r02_unet_traffic_summary.txt
SEMABS_UNET_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv|gn_apply|fold|maxpool|ncdhw' --launch-skip 0 -c 400 --csv --log-file gpurun_out/r02_launch_unet.csv python tools/profile_step.py unet 1 > gpurun_out/ncu_ev2_launch.log 2>&1
FILT='sm__ops_path\|sm__mem_tensor_cycles_active\.\(max\|min\|sum\)\|pipe_tensor_cycles_active\.\(max\|min\|sum\)\| 0 \| 0$\|TriageCompute\|device__attribute'
cap() { local name=$1 rx=$2 skip=$3 cnt=$4 top=$5; shift 5
  SEMABS_UNET_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu_ev2_$name.log 2>&1
  { echo "# ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c $cnt $*"; python tools/ncu_summary.py /tmp/$name.ncu-rep 2>&1 | grep -v "$FILT"; python tools/ncu_hot.py /tmp/$name.ncu-rep $top 2>&1; } > gpurun_out/r02_ncu_full_$name.txt
  echo "$name: $(grep -c '^== ' gpurun_out/r02_ncu_full_$name.txt) sections"; }
cap convt 'convt_allparity' 0 2 12 python tools/profile_step.py unet 0
cap halo_fused 'conv3d_halo_pair' 8 3 12 python tools/profile_step.py unet 0
cap igemm64 'conv3d_igemm_kernel<64, 64' 0 2 10 python tools/profile_step.py unet 0
timeout 1500 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; tail -2 gpurun_out/r02_bench_line.err
timeout 1500 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2> gpurun_out/r02_bench_reference_line.err
timeout 120 tools/ubench/tmem_bw > gpurun_out/r02_ubench_tmem.txt 2>&1
timeout 120 tools/ubench/mma_issue > gpurun_out/r02_ubench_mma_issue.txt 2>&1
timeout 300 python tools/attn_trace.py 2>&1 | grep -v Warn > gpurun_out/r02_attn_trace.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_line.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'ms/step', d['ms_per_step'])
r=d['roofline']; print('gemm frac', r['frac'], 'whole', r['whole_path_frac'])
for k in ('voxel','pipeline','train','cuda_eager','faithful_ours','cpu_baseline'):
    v=d.get(k); print(k, json.dumps(v)[:500] if v else None)
print(open('gpurun_out/r02_bench_reference_line.json').read()[:600])
PY
