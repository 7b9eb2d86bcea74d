#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NPROC:-2} --master-addr 127.0.0.1 --master-port 29511 tools/check_grad_buckets.py > gpurun_out/r02_grad_buckets_${NPROC:-2}gpu.json 2> gpurun_out/r02_grad_buckets_${NPROC:-2}gpu.err
echo rc=$?; cat gpurun_out/r02_grad_buckets_${NPROC:-2}gpu.json; tail -5 gpurun_out/r02_grad_buckets_${NPROC:-2}gpu.err
