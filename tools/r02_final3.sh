#!/bin/bash
# Round-2 closing evidence at HEAD: GPU suite, both bench arms.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -2 > gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; tail -1 gpurun_out/r02_bench_line.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2> gpurun_out/r02_bench_reference_line.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_line.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'ms/step', d['ms_per_step'])
r=d['roofline']; print('gemm frac', r['frac'], 'whole', r['whole_path_frac'], 'traffic', r['traffic'])
v=d['voxel']; print('voxel', v['value'], v['ms_per_step'], v['e2e']['value'], v['roofline']['frac'], v['roofline']['traffic'])
print('train', d['train']['ms_per_step'], 'pipe', d['pipeline']['value'], 'eager', d['cuda_eager']['relevancy']['value'], d['cuda_eager']['voxel']['fp32']['value'])
for k in r['kernels'][:8]: print(k)
print(open('gpurun_out/r02_bench_reference_line.json').read()[:300])
PY
