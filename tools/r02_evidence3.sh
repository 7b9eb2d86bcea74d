#!/bin/bash
# Round-2 evidence, third part (after the fully folded level 0): GPU suite, UNet traffic + launch list, both bench arms.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -3 > gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_gpu_tests.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
SEMABS_UNET_GRAPH=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_unet_traffic.csv python tools/profile_step.py unet 1 > gpurun_out/ncu_ev3_traffic.log 2>&1
python tools/summarize_traffic.py gpurun_out/r02_unet_traffic.csv > gpurun_out/r02_unet_traffic_summary.txt; head -14 gpurun_out/r02_unet_traffic_summary.txt
SEMABS_UNET_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv|gn_apply|fold|maxpool|ncdhw' --launch-skip 0 -c 400 --csv --log-file gpurun_out/r02_launch_unet.csv python tools/profile_step.py unet 1 > gpurun_out/ncu_ev3_launch.log 2>&1
timeout 1500 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; tail -2 gpurun_out/r02_bench_line.err
timeout 1500 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2> gpurun_out/r02_bench_reference_line.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_line.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'ms/step', d['ms_per_step'])
r=d['roofline']; print('gemm frac', r['frac'], 'whole', r['whole_path_frac'])
v=d['voxel']; print('voxel', v['value'], v['ms_per_step'], v['e2e']['value'], v['roofline']['frac'], v['roofline']['traffic'])
print('train', d['train']['ms_per_step'], 'pipe', d['pipeline']['value'], 'eager', d['cuda_eager']['relevancy']['value'], d['cuda_eager']['voxel']['fp32']['value'])
PY
