"""Debug aid: relevance-only vs full call of the attention backward generations over many units (prints where they differ)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semabs_b200 import ops
dev = "cuda"
for (B, T, H, P) in [(29, 257, 16, 5), (64, 257, 16, 1), (29, 257, 16, 1), (64, 257, 16, 5)]:
    g = torch.Generator(device=dev).manual_seed(B * T + P)
    d = H * 64
    qkv = torch.randn(B * T, 3 * d, device=dev, generator=g)
    qkv[:, :d] *= 0.125 * 1.5
    Tp = (T + 15) // 16 * 16
    probs16 = torch.empty(B * H, T, Tp, device=dev, dtype=torch.float16)
    o32 = torch.empty(B * T, d, device=dev)
    ops.attn_fwd(qkv, B=B, T=T, H=H, probs=None, probs16=probs16, o32=o32)
    qkv16 = qkv.half()
    dO = torch.randn(P * B * T, d, device=dev, generator=g).half()
    r = torch.rand(P * B, T, device=dev, generator=g)
    for gen in (2, 3):
        for rep in range(2):
            delta = torch.empty(P * B * H, T, device=dev)
            wpart = torch.full((P * B * H, T), float("nan"), device=dev)
            dqkv16 = torch.full((P * B * T, 2 * 3 * d), float("nan"), device=dev, dtype=torch.float16)
            ops.attn_bwd_tc(qkv16, probs16, o32, dO, d, r, delta, wpart, dqkv16, P=P, B=B, T=T, H=H, splits=2, positive_only=True, generation=gen)
            w2 = torch.full_like(wpart, float("nan"))
            ops.attn_bwd_tc(qkv16, probs16, o32, dO, d, r, delta, w2, None, P=P, B=B, T=T, H=H, splits=2, positive_only=True, need_dqkv=False, generation=gen)
            torch.cuda.synchronize()
            bad = (w2 != wpart)
            idx = bad.nonzero()
            print(f"B={B} T={T} H={H} P={P} gen {gen} rep {rep}: mismatches {int(bad.sum())} of {bad.numel()}, nan full {int(wpart.isnan().sum())} rel-only {int(w2.isnan().sum())}",
                  f"maxdiff {(w2 - wpart).abs().nan_to_num(9e9).max().item():.3e}", idx[:6].tolist(), flush=True)
