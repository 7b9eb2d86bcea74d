"""Debug helper: per-unit comparison of the UNet backward intermediates (dz, dxn) against autograd through the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import unet_oracle
from semabs_b200.unet3d import ResidualUNet3D
from semabs_b200 import unet3d_bwd, ops

dev = "cuda"
torch.manual_seed(0)
m = ResidualUNet3D(in_channels=16, out_channels=16, f_maps=16, num_groups=8, num_levels=3).to(dev)
g = torch.Generator().manual_seed(1)
shape = (16, 16, 16); N = 2
x = torch.randn(N, 16, *shape, generator=g); gy = torch.randn(N, 16, *shape, generator=g)
sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
rec = []
orig = F.conv3d
def conv(inp, w, *a, **k):
    out = orig(inp, w, *a, **k)
    if w.shape[-1] == 3:
        inp.retain_grad(); out.retain_grad(); rec.append((inp, w, out))
    return out
F.conv3d = conv
xo = x.clone().requires_grad_(True)
yo = unet_oracle.residual_unet3d(sd, xo); yo.backward(gy)
F.conv3d = orig
# oracle conv order: enc0 c1,c2,c3, enc1 ..., enc2 ..., dec0 c1.., dec1 c1..
names = [f"{p}.{j}" for p in ("enc0", "enc1", "enc2", "dec0", "dec1") for j in (1, 2, 3)]
ref = {n: (r[2].grad, r[0].grad) for n, r in zip(names, rec)}  # (dz, dxn) NCDHW

mine = {}
orig_unit = unet3d_bwd.UNetBackward._unit
def unit(self, pk, fpk, prefix, j, sc, g_, mask, x_raw, x_stats, **kw):
    orig_unit(self, pk, fpk, prefix, j, sc, g_, mask, x_raw, x_stats, **kw)
    torch.cuda.synchronize()
    Nn, dims, c_out, c_in_pad = kw["N"], kw["dims"], kw["c_out"], kw["c_in_pad"]
    S = dims[0] * dims[1] * dims[2]
    sp = 2
    dz = self.unet._ws[("bwd_dz_op", (Nn * S * sp * c_out,), torch.float16, str(kw["dev"]))].view(Nn, S, sp, c_out).float().sum(2)
    dxn = self.unet._ws[("bwd_dxn", (Nn * S * c_in_pad,), torch.float32, str(kw["dev"]))].view(Nn, S, c_in_pad).clone()
    scale = self._scales[self._slot_i - 1 if False else 0]
    mine[f"{prefix}.{j}"] = (dz.clone(), dxn, None, kw["dx"].clone() if kw.get("dx") is not None else None, g_.t.clone().view(Nn, S, c_out),
                              (1.0 / g_.scale.item()) if g_.scale is not None else 1.0)
unet3d_bwd.UNetBackward._unit = unit
xg = x.to(dev).requires_grad_(True)
y = m(xg); y.backward(gy.to(dev))
def r2(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
def cl(t): return t.permute(0, 2, 3, 4, 1).reshape(t.shape[0], -1, t.shape[1])
for n in ["dec1.3", "dec1.2", "dec1.1", "dec0.3", "dec0.2", "dec0.1", "enc2.3", "enc2.2", "enc2.1", "enc1.3", "enc1.2", "enc1.1", "enc0.3", "enc0.2", "enc0.1"]:
    dz, dxn, _, dx, gin, ginv = mine[n]
    rdz, rdxn = cl(ref[n][0]), cl(ref[n][1])
    # scales unknown here: fit the best scalar
    a = (dz.cpu() * rdz).sum() / (dz.cpu() ** 2).sum()
    b = (dxn.cpu() * rdxn).sum() / (dxn.cpu() ** 2).sum()
    e_dz = r2(dz.cpu() * a, rdz); e_dxn = r2(dxn.cpu() * b, rdxn)
    d = (dz.cpu() * a - rdz).abs()
    print(f"{n}: dz err {e_dz:.2e} (scale {1/a:.3g}), dxn err {e_dxn:.2e}; dz max abs err {d.max():.3g} vs rms {rdz.pow(2).mean().sqrt():.3g}; n(|err|>0.01 rms) = {(d > 0.01 * rdz.pow(2).mean().sqrt()).sum().item()}")
print("input grad err", r2(xg.grad.cpu(), xo.grad))
