"""Timeline of one mid-grid CTA of the tcgen05 attention forward (semabs_debug_attn_fwd_trace)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semabs_b200 import ops
from semabs_b200._lib import lib

dev = torch.device("cuda", 0)
B, T, H = 95, int(os.environ.get("T", "257")), 16
d = H * 64
splits = int(os.environ.get("SPLITS", "1"))
qkv = torch.randn(B * T, 3 * d, device=dev)
qkv[:, :d] *= 0.25
qkv16 = ops.split_f16(qkv) if splits == 2 else qkv.half()
ldp = (T + 15) // 16 * 16
probs16 = torch.empty(B * H, T, ldp, device=dev, dtype=torch.float16)
o32 = torch.empty(B * T, d, device=dev)
o16 = torch.empty(B * T, 2 * d, device=dev, dtype=torch.float16)
tr = torch.zeros(64, dtype=torch.int64, device=dev)
run = lambda: ops.attn_fwd_tc(qkv16, in_splits=splits, B=B, T=T, H=H, probs16=probs16, o32=o32, o16=o16, o_splits=2, causal=False)
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
print(f"T={T} in_splits={splits}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch ({B * H} CTAs)")
lib().semabs_debug_attn_fwd_trace(ops.ptr(tr))
run(); torch.cuda.synchronize()
lib().semabs_debug_attn_fwd_trace(None)
t = tr.cpu().tolist()
t0 = t[0]
names = {1: "wait S", 2: "S ready", 3: "max done", 4: "exp done", 5: "packed (P ready)", 6: "O ready", 7: "stored"}
print("start 0")
for mt in range(3):
    if t[1 + 8 * mt]:
        print(f"tile {mt}: " + ", ".join(f"{names[k]} {t[k + 8 * mt] - t0}" for k in range(1, 8)))
for w in range(3):
    if t[32 + 4 * w]:
        print(f"tail warp {w + 1}: start {t[32 + 4 * w] - t0}, K/V ready {t[33 + 4 * w] - t0}, row done {t[34 + 4 * w] - t0}")
print("TMEM freed", t[31] - t0)
