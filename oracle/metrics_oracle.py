"""TEST INFRASTRUCTURE ONLY — CPU restatement of utils.prediction_analysis (utils.py:338-380) and utils.voxelize_points
(utils.py:617-665; torch_scatter.scatter(reduce="max") restated as scatter_reduce amax with empty cells = 0, like
oracle/ref_import.py).  Pinned: `python -m oracle.metrics_oracle` imports the unmodified reference `utils` (shims of
oracle/ref_import.py) and asserts equality on seeded inputs — last run in the build container: identical (NaNs included)."""
from __future__ import annotations

import numpy as np
import torch

from .unet_oracle import point_grid_indices


def prediction_analysis(prediction, label, ignore):
    stats = {"precision": [], "recall": [], "false_negative": [], "false_positive": [], "iou": []}
    for b in range(ignore.shape[0]):
        for p in range(ignore.shape[1]):
            mask = ~ignore.bool()[b, p]
            l, q = label.bool()[b, p][mask], prediction.bool()[b, p][mask]
            tp = torch.logical_and(l, q).float().sum().item()
            stats["iou"].append((torch.logical_and(q, l).sum().float() / torch.logical_or(q, l).sum().float()).item())
            stats["precision"].append(tp / q.float().sum().item() if q.float().sum().item() != 0 else np.nan)
            stats["recall"].append(tp / l.float().sum().item() if l.float().sum().item() != 0 else np.nan)
            stats["false_negative"].append(torch.logical_and(l, ~q).float().mean().item())
            stats["false_positive"].append(torch.logical_and(~l, q).float().mean().item())
    return stats


def _scatter_max(xyz, feat, bounds, grid_shape):
    idx = point_grid_indices(xyz, bounds, grid_shape)
    flat = (idx[..., 0] * grid_shape[1] + idx[..., 1]) * grid_shape[2] + idx[..., 2]
    n_vox = int(np.prod(grid_shape))
    return torch.zeros(feat.shape[0], n_vox).scatter_reduce(1, flat, feat, reduce="amax", include_self=False)


def voxelize_points(prediction, label, xyz_pts, voxel_shape, scene_bounds, ignore_pts):
    B, P, n = prediction.shape
    xyz = xyz_pts.reshape(B * P, n, 3)
    vp = _scatter_max(xyz, prediction.float().view(B * P, n), scene_bounds, voxel_shape)
    vl = _scatter_max(xyz, (label.float().view(B * P, n) - 0.5) * 2, scene_bounds, voxel_shape)
    missing = vl == 0.0
    vi = _scatter_max(xyz, ignore_pts.float().view(B * P, n), scene_bounds, voxel_shape).bool() | missing
    return {"prediction": (vp > 0).view(B, P, -1), "label": (vl > 0).float().view(B, P, -1), "ignore": vi.view(B, P, -1)}


def _pin():
    """utils.py cannot be imported whole here (dataset.py / fusion.py pull in a dozen absent packages), so the three
    functions are compiled from the unmodified source text of /root/reference/utils.py and run against the reference's own
    VirtualGrid (net.py) with the torch_scatter restatement of oracle/ref_import.py."""
    import ast, importlib, sys, types
    from typing import Tuple

    import filelock, pandas  # noqa: F401  (the real modules must be in sys.modules before the stubs are installed)

    from . import ref_import

    ref_import.install_shims()
    if ref_import.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_import.REF_ROOT)
    net = importlib.import_module("net")
    src = open(ref_import.REF_ROOT + "/utils.py").read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("iou", "prediction_analysis", "voxelize_points")]
    assert len(wanted) == 3
    for n in wanted:
        n.decorator_list = []  # @typechecked (typeguard 2 syntax)
    ns = {"torch": torch, "np": np, "Tuple": Tuple, "VirtualGrid": net.VirtualGrid, "TensorType": sys.modules["torchtyping"].TensorType}
    if not hasattr(np, "NAN"):
        np.NAN = np.nan
    exec(compile(ast.Module(body=wanted, type_ignores=[]), "utils.py", "exec"), ns)
    utils = types.SimpleNamespace(**{k: ns[k] for k in ("prediction_analysis", "voxelize_points")})
    g = torch.Generator().manual_seed(0)
    B, P, n = 2, 3, 4000
    pred = torch.rand(B, P, n, generator=g) < 0.3
    lab = torch.rand(B, P, n, generator=g) < 0.2
    ign = torch.rand(B, P, n, generator=g) < 0.25
    pred[0, 1] = False          # precision NaN
    lab[1, 0] = False           # recall NaN
    ign[1, 2] = True            # everything ignored: all NaN
    bounds = ((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9))
    xyz = torch.tensor(bounds[0]) + (torch.tensor(bounds[1]) - torch.tensor(bounds[0])) * (torch.rand(B, P, n, 3, generator=g) * 1.1 - 0.05)
    a, b = utils.prediction_analysis(prediction=pred, label=lab, ignore=ign), prediction_analysis(pred, lab, ign)
    for k in a:
        assert np.allclose(a[k], b[k], equal_nan=True, rtol=0, atol=0), k
    va = utils.voxelize_points(prediction=pred, label=lab, xyz_pts=xyz, voxel_shape=(8, 8, 8), scene_bounds=torch.tensor(bounds),
                               ignore_pts=ign, device="cpu")
    vb = voxelize_points(pred, lab, xyz, (8, 8, 8), bounds, ign)
    for k in va:
        assert torch.equal(va[k].float(), vb[k].float()), k
    print("metrics oracle pinned to /root/reference/utils.py: identical")


if __name__ == "__main__":
    _pin()
