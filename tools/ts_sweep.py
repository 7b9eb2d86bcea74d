"""Descriptor sweep for semabs_selftest_ts_mma (development aid): prints the error of each (lbo, sbo) candidate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semabs_b200 import ops
g = torch.Generator(device="cuda").manual_seed(5)
for Kd in (16, 64):
    A = torch.randn(128, Kd, device="cuda", generator=g).half()
    Bm = torch.randn(Kd, 64, device="cuda", generator=g).half()
    ref = (A.double() @ Bm.double()).float()
    for lbo, sbo in [(16, 1024), (1024, 16), (0, 1024), (1024, 1024), (128, 1024), (1024, 128), (2048, 1024), (1024, 2048), (8192, 1024)]:
        D = torch.full((128, 64), float("nan"), device="cuda")
        try:
            ops.selftest_ts_mma(A, Bm, D, lbo, sbo)
            torch.cuda.synchronize()
            print(f"Kd={Kd} lbo={lbo} sbo={sbo}: max err {(D - ref).abs().max().item():.4g}  ref max {ref.abs().max().item():.3g}", flush=True)
        except Exception as e:
            print(f"Kd={Kd} lbo={lbo} sbo={sbo}: FAILED {e}", flush=True)
            sys.exit(1)
