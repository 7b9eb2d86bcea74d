"""Run the same SemAbs3D loss + backward several times on one GPU, snapshot every tensor argument of every
semabs_b200.ops call, and report the first calls whose tensors differ between runs by more than atomics noise."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semabs_b200 import ops, train  # noqa: E402
from semabs_b200.net import SemAbs3D  # noqa: E402

BOUNDS = ((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9))
dev = torch.device("cuda", 0)
LOG = []
THRESH = float(os.environ.get('THRESH', '3e-4'))


def wrap(name, fn):
    def inner(*a, **k):
        r = fn(*a, **k)
        snap = {}
        for i, v in enumerate(a):
            if torch.is_tensor(v):
                snap[f"arg{i}"] = v.detach().clone()
        for kk, v in k.items():
            if torch.is_tensor(v):
                snap[kk] = v.detach().clone()
        info = {kk: v for kk, v in k.items() if isinstance(v, (int, float, bool))}
        LOG.append((name, snap, info))
        return r
    return inner


for n in dir(ops):
    f = getattr(ops, n)
    if isinstance(f, types.FunctionType) and f.__module__ == ops.__name__ and not n.startswith("_"):
        setattr(ops, n, wrap(n, f))


def run():
    LOG.clear()
    torch.manual_seed(5)
    m = SemAbs3D(voxel_shape=(32, 32, 32), scene_bounds=BOUNDS, unet_num_channels=16, unet_f_maps=16, unet_num_groups=8,
                 unet_num_levels=4, network_inputs=["saliency"], use_pts_feat_extractor=True, pts_feat_extractor_hidden_dim=128,
                 reduce_method="max", device=str(dev), batch_size=1).to(dev)
    g = torch.Generator().manual_seed(50)
    lo, hi = torch.tensor(BOUNDS[0]), torch.tensor(BOUNDS[1])
    B, P, n_in, n_out = 1, 2, 4000, 6000
    batch = dict(input_xyz_pts=(lo + (hi - lo) * torch.rand(B, n_in, 3, generator=g)).to(dev),
                 input_feature_pts=torch.randn(B, P, n_in, 1, generator=g).to(dev), tsdf_vol=torch.ones(B, 1, device=dev),
                 output_xyz_pts=(lo + (hi - lo) * torch.rand(B, P, n_out, 3, generator=g)).to(dev),
                 output_label_pts=(torch.rand(B, P, n_out, generator=g) < 0.15).float().to(dev),
                 out_of_bounds_pts=torch.zeros(B, P, n_out, dtype=torch.bool, device=dev),
                 out_of_frustum_pts_mask=torch.zeros(B, P, n_out, dtype=torch.bool, device=dev), patch_labels=[("a",), ("b",)])
    stats, _ = train.get_losses_ovssc(m, batch)
    stats["loss"].backward()
    torch.cuda.synchronize()
    out = list(LOG)
    LOG.clear()
    return out, {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}


def rel(a, b):
    a, b = a.double(), b.double()
    if not torch.isfinite(a).all() or not torch.isfinite(b).all():
        return float("nan")
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


ref_log, ref_g = run()
for r in range(1, int(os.environ.get("RUNS", "5"))):
    log, g = run()
    assert len(log) == len(ref_log)
    worst = max(rel(g[k], ref_g[k]) for k in g)
    shown = 0
    for i, ((n0, s0, info), (n1, s1, _)) in enumerate(zip(ref_log, log)):
        for k in s0:
            if s0[k].shape != s1[k].shape:
                continue
            if not s0[k].is_floating_point():
                if not torch.equal(s0[k], s1[k]) and worst > 3e-4 and shown < 30 and s0[k].numel() <= 4:
                    print(f"run {r}: call {i} {n0}.{k} int tensor differs: {s0[k].tolist()} vs {s1[k].tolist()}")
                continue
            d = rel(s1[k], s0[k])
            if k in ("mask", "add_mask"):
                flips = ((s1[k] > 0) != (s0[k] > 0)).sum().item()
                if flips and worst > 3e-4 and shown < 30:
                    print(f"run {r}: call {i} {n0}.{k} ReLU mask flips {flips} of {s0[k].numel()} (rel diff of the mask tensor {d:.2e})")
            if d > THRESH or d != d:
                if shown < 30 and worst > 3e-4 and i > 73:
                    print(f"run {r}: call {i} {n0}.{k} shape {tuple(s0[k].shape)} rel diff {d:.3e} max|d| {(s1[k].double()-s0[k].double()).abs().max().item():.3e} "
                          f"n_diff {(s1[k] != s0[k]).sum().item()} {info if shown < 3 else ''}")
                shown += 1
    print(f"run {r}: worst gradient rel diff {worst:.3e}; tensors over {THRESH}: {shown}")
    del log, g
