"""Timeline of CTA 0 of the chunk-pipelined attention backward (semabs_debug_attn_trace): cycles between pipeline events per item."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from semabs_b200 import ops
from semabs_b200._lib import lib
dev = "cuda"
B, T, H, P = int(os.environ.get("TR_B", "95")), 257, 16, 16
g = torch.Generator(device=dev).manual_seed(1)
d = H * 64
qkv = torch.randn(B * T, 3 * d, device=dev, generator=g); qkv[:, :d] *= 0.19
Tp = (T + 15) // 16 * 16
probs16 = torch.empty(B * H, T, Tp, device=dev, dtype=torch.float16)
o32 = torch.empty(B * T, d, device=dev)
ops.attn_fwd(qkv, B=B, T=T, H=H, probs=None, probs16=probs16, o32=o32)
qkv16 = qkv.half()
dO = torch.randn(P * B * T, d, device=dev, generator=g).half()
r = torch.rand(P * B, T, device=dev, generator=g)
delta = torch.empty(P * B * H, T, device=dev); wpart = torch.empty(P * B * H, T, device=dev)
dqkv16 = torch.empty(P * B * T, 3 * d, device=dev, dtype=torch.float16)
tr = torch.zeros(2, 16, 64, dtype=torch.int64, device=dev)
for it in range(2):
    lib().semabs_debug_attn_trace(ops.ptr(tr) if it == 1 else None)
    ops.attn_bwd_tc(qkv16, probs16, o32, dO, d, r, delta, wpart, dqkv16, P=P, B=B, T=T, H=H, splits=1, positive_only=True)
    torch.cuda.synchronize()
lib().semabs_debug_attn_trace(None)
t = tr.cpu().numpy()
names = {0: "ctl p0 seen", 1: "ctl acc free", 2: "ctl issued c0", 3: "ctl p1 seen", 4: "ctl issued c1", 5: "h0 top", 6: "h0 G seen", 7: "h0 arrived",
         8: "h0 epi done", 10: "h1 top", 11: "h1 G seen", 12: "h1 arrived", 13: "h1 epi done", 15: "ctl MMAs done"}
for k, kn in enumerate(["row pass", "column pass"]):
    t0 = t[k][np.nonzero(t[k])].min()
    print(f"== {kn}: cycles since the first event, items 16..27 of CTA 0 (one column per item)")
    for e in sorted(names):
        if t[k, e].any():
            print(f"  {names[e]:15s}", " ".join(f"{int(v - t0):7d}" for v in t[k, e, 16:28]))
    per = np.diff(t[k, 0, 8:60]).astype(float)
    print(f"  period per item (ctl p0 seen): median {np.median(per):.0f}, mean {per.mean():.0f} cycles")
