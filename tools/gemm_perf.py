"""Quick device-side timing of semabs_gemm_f16 at ViT-L/14 shapes (CUDA events, L2-flushed between reps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semabs_b200 import ops

def bench(M, N, K, splits=1, reps=10):
    a = torch.randn(M, splits * K, device="cuda").half()
    b = torch.randn(N, K, device="cuda").half()
    out = torch.empty(M, N, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.gemm_f16(a, b, a_splits=splits, out_f32=out)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_f16(a, b, a_splits=splits, out_f32=out); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"M={M} N={N} K={K} splits={splits}: {t*1e3:.1f} us  {2*M*N*K*splits/t/1e9:.1f} TFLOP/s", flush=True)

if __name__ == "__main__":
    for shp in [(8224, 3072, 1024), (8224, 4096, 1024), (8224, 1024, 4096), (131584, 4096, 1024), (131584, 1024, 4096), (131584, 1024, 3072)]:
        bench(*shp)
    bench(8224, 3072, 1024, splits=2)
