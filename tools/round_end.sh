#!/bin/bash
# Round-end evidence run on a B200 box (gpurun): GPU test-suite, both bench arms, the ncu launch list of the bench command
# and the DRAM-traffic / full-set captures quoted in DESIGN.md. Outputs land in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/tests.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
# one image of the bench workload: 3 tile batches of 95 (the first engine call is the warm-up)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --images 1 --skip-cpu --skip-train --skip-pipeline > gpurun_out/b_ncu.log 2>&1
VIT_B=95 ncu --metrics $M --clock-control none -k regex:gemm_f16_tn --launch-skip 100 -c 100 --csv --log-file gpurun_out/gemm_traffic.csv \
    python tools/profile_step.py vit 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm --launch-skip 150 -c 3 -f -o gpurun_out/r01_igemm_fused \
    python tools/profile_step.py unet 1 > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3d_wgrad --launch-skip 84 -c 2 -f -o gpurun_out/r01_wgrad_l0 \
    python tools/profile_train.py 4 1 > gpurun_out/ncu3.log 2>&1
cat gpurun_out/tests.log
cat gpurun_out/bench.json | cut -c1-4000
