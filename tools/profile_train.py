"""One VOOL train step (bench.py:bench_train workload) for ncu launch lists: `ncu ... python tools/profile_train.py [descs]`."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

descs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.cuda.set_device(0)
r = bench.bench_train(torch.device("cuda", 0), 0, 1, bench.peaks(), num_descs=descs, steps=steps, warmup=1)
print(json.dumps(r))
