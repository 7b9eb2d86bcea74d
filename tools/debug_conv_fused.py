"""Debug helper: precise igemm conv (fused two-stage schedule) on the shapes of the VOOL training test, repeated for
determinism, against torch fp32 convs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from semabs_b200 import ops
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"

def cl(x):
    xc = x.permute(0, 2, 3, 4, 1).contiguous(); hi = xc.half()
    return torch.cat([hi, (xc - hi.float()).half()], dim=-1).contiguous()
def pack(w2d):
    hi = w2d.half(); return torch.cat([hi, (w2d - hi.float()).half()], dim=1).contiguous()

for (N, D, Ci, Co) in [(3, 4, 64, 64), (3, 8, 32, 32), (3, 16, 16, 16), (2, 4, 64, 64), (3, 4, 64, 32), (3, 8, 32, 16), (1, 4, 64, 64), (5, 4, 128, 64)]:
    g = torch.Generator(device=dev).manual_seed(N * 100 + D + Ci)
    x = torch.randn(N, Ci, D, D, D, device=dev, generator=g)
    w = torch.randn(Co, Ci, 3, 3, 3, device=dev, generator=g) / (27 * Ci) ** 0.5
    ref = F.conv3d(x, w, padding=1).permute(0, 2, 3, 4, 1)
    outs = []
    for rep in range(4):
        out = torch.full((N, D, D, D, Co), float("nan"), device=dev)
        ops.conv3d(cl(x), pack(w.permute(0, 2, 3, 4, 1).reshape(Co, -1)), kind=ops.CONV_3X3X3, N=N, D=D, H=D, W=D, C_in=Ci, C_out=Co,
                   a_splits=2, w_splits=2, precise=True, out32=out)
        torch.cuda.synchronize()
        outs.append(out)
    err = [((o - ref).abs().max() / ref.abs().max()).item() for o in outs]
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    per_n = [((outs[-1][n] - ref[n]).abs().max() / ref.abs().max()).item() for n in range(N)]
    print(f"N={N} D={D} Ci={Ci} Co={Co}: err {max(err):.2e} deterministic={same} per-sample {['%.1e' % e for e in per_n]}")

print("--- kind 3 (adjoint of the transposed conv), precise, chained parity launches")
for (N, D, Ci_T, Co_T) in [(3, 4, 64, 32), (2, 4, 64, 32), (3, 8, 32, 16), (1, 4, 64, 32), (3, 4, 128, 64)]:
    g = torch.Generator(device=dev).manual_seed(N + D + Ci_T)
    w = torch.randn(Ci_T, Co_T, 3, 3, 3, device=dev, generator=g) / (8 * Ci_T) ** 0.5   # ConvTranspose3d weight
    dy = torch.randn(N, Co_T, 2 * D, 2 * D, 2 * D, device=dev, generator=g)
    xr = torch.zeros(N, Ci_T, D, D, D, device=dev, requires_grad=True)
    F.conv_transpose3d(xr, w, stride=2, padding=1, output_padding=1).backward(dy)
    ref = xr.grad.permute(0, 2, 3, 4, 1)
    wa = pack(w.permute(0, 2, 3, 4, 1).reshape(Ci_T, -1))
    outs = []
    for rep in range(4):
        dx = torch.full((N, D, D, D, Ci_T), float("nan"), device=dev)
        for q in range(8):
            ops.conv3d(cl(dy), wa, kind=ops.CONV_TRANSPOSE_ADJOINT, parity=q, N=N, D=D, H=D, W=D, C_in=Co_T, C_out=Ci_T, a_splits=2,
                       w_splits=2, precise=True, out32=dx, residual=dx if q > 0 else None)
        torch.cuda.synchronize()
        outs.append(dx)
    err = [((o - ref).abs().max() / ref.abs().max()).item() for o in outs]
    print(f"N={N} D={D} Ci_T={Ci_T} Co_T={Co_T}: err {max(err):.2e} deterministic={all(torch.equal(outs[0], o) for o in outs[1:])}")
