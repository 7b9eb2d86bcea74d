#!/bin/bash
# final round-2 lines: both bench arms as the driver runs them, the GPU suite, micro-benchmarks and the attention pipeline trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -3 > gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_gpu_tests.log
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
    v=d.get(k); print(k, json.dumps(v)[:600] if v else None)
print(open('gpurun_out/r02_bench_reference_line.json').read()[:800])
PY
