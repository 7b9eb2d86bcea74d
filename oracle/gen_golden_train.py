"""TEST INFRASTRUCTURE ONLY — golden vectors for the training-side boundary, produced by running
the UNMODIFIED reference (`train_ovssc.get_losses`, `train_vool.get_losses`, `utils.config_parser`, `utils.get_net`'s
scheduler, imported from /root/reference through oracle/ref_import.py) on CPU:

    python -m oracle.gen_golden_train        ->  tests/golden/train_golden.json

`get_losses` is driven by a stub network that returns fixed seeded logits (so the fixture pins the loss / accuracy /
per-cutoff IoU table independently of the UNet), on a batch that exercises padding patches, out-of-bounds and
out-of-frustum points, an empty-prediction patch (NaN precision) and `balance_positive_negative`.  The tests regenerate
the inputs from `make_ovssc_case` / `make_vool_case` below (pure torch, no reference needed)."""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BOUNDS = [[-1.0, -1.0, -0.1], [1.0, 1.0, 1.9]]
CUTOFFS_OVSSC = [-1.0, 0.0, 0.7]
CUTOFFS_VOOL = [-2.0, -0.5]


class StubNet:
    """net(**batch) -> fixed logits; `device` like the reference modules expose"""

    def __init__(self, logits, device="cpu"):
        self.logits, self.device = logits, device

    def __call__(self, **batch):
        return self.logits


def _pts(g, *lead):
    lo, hi = torch.tensor(BOUNDS[0]), torch.tensor(BOUNDS[1])
    return lo + (hi - lo) * (torch.rand(*lead, 3, generator=g) * 1.1 - 0.05)


def make_ovssc_case(seed=0, B=2, P=3, n=6000):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, P, n, generator=g) * 2
    logits[1, 0] = -5.0  # a patch with no positive prediction at any cutoff: precision NaN
    batch = dict(output_xyz_pts=_pts(g, B, P, n), output_label_pts=(torch.rand(B, P, n, generator=g) < 0.2).float(),
                 out_of_bounds_pts=torch.rand(B, P, n, generator=g) < 0.1, out_of_frustum_pts_mask=torch.rand(B, P, n, generator=g) < 0.1,
                 patch_labels=[("chair", "table"), ("sofa", ""), ("", "lamp")], scene_id=["scene_a", "scene_b"],
                 scene_bounds=torch.tensor(BOUNDS))
    return logits, batch


def make_vool_case(seed=1, B=2, D=3, n=5000):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, D, n, generator=g) * 2 - 1
    batch = dict(output_xyz_pts=_pts(g, B, D, n), output_label_pts=(torch.rand(B, D, n, generator=g) < 0.15).float(),
                 out_of_bounds_pts=torch.rand(B, D, n, generator=g) < 0.1,
                 spatial_relation_name=[("behind", "on"), ("in", "[pad]"), ("[pad]", "on the left of")],
                 target_obj_name=[("mug", "book"), ("pen", ""), ("", "plant")], reference_obj_name=[("desk", "shelf"), ("cup", ""), ("", "sofa")],
                 scene_id=["scene_c", "scene_d"], scene_bounds=torch.tensor(BOUNDS))
    return logits, batch


def _jsonable(x):
    if isinstance(x, float) and math.isnan(x):
        return None
    if isinstance(x, (np.floating, np.integer)):
        return _jsonable(x.item())
    if torch.is_tensor(x):
        return _jsonable(x.item())
    return x


def frame_to_dict(df):
    return {"columns": list(df.columns), "rows": [[_jsonable(v) for v in row] for row in df.itertuples(index=False, name=None)]}


def main():
    import filelock, pandas, transformers  # noqa: F401  (real modules first; ref_import only stubs what is absent)
    from transformers import get_scheduler  # noqa: F401

    from oracle import ref_import

    ref_import.install_shims()
    t3d = sys.modules["transforms3d"]
    for sub in ("affines", "euler", "quaternions"):
        setattr(t3d, sub, ref_import._stub("transforms3d." + sub))
    ref_ovssc = ref_import.import_reference_module("train_ovssc")
    ref_vool = ref_import.import_reference_module("train_vool")
    ref_utils = ref_import.import_reference_module("utils")
    # train_ovssc.get_detailed_stats calls utils.voxelize_points without `device` (default "cuda", train_ovssc.py:40-47)
    orig_vox = ref_utils.voxelize_points
    ref_utils.voxelize_points = lambda *a, **k: orig_vox(*a, **{**k, "device": "cpu"})
    out = {}
    args = ref_utils.config_parser().parse_args(["--file_path", "x"])
    out["config_defaults"] = {k: (v if not isinstance(v, torch.Tensor) else v.tolist()) for k, v in vars(args).items()}

    for bal in (False, True):
        logits, batch = make_ovssc_case()
        sb = batch.pop("scene_bounds")
        logits = logits.requires_grad_(True)
        stats, df = ref_ovssc.get_losses(StubNet(logits), batch, cutoffs=CUTOFFS_OVSSC, balance_positive_negative=bal, scene_bounds=sb)
        stats["loss"].backward()
        out[f"ovssc_bal{int(bal)}"] = {"stats": {k: _jsonable(float(v)) for k, v in stats.items()}, "frame": frame_to_dict(df),
                                       "dlogits_sum_abs": float(logits.grad.abs().sum()), "dlogits_head": logits.grad.flatten()[:8].tolist()}
        logits, batch = make_vool_case()
        sb = batch.pop("scene_bounds")
        logits = logits.requires_grad_(True)
        stats, df = ref_vool.get_losses(StubNet(logits), batch, cutoffs=CUTOFFS_VOOL, balance_positive_negative=bal, scene_bounds=sb)
        stats["loss"].backward()
        out[f"vool_bal{int(bal)}"] = {"stats": {k: _jsonable(float(v)) for k, v in stats.items()}, "frame": frame_to_dict(df),
                                      "dlogits_sum_abs": float(logits.grad.abs().sum()), "dlogits_head": logits.grad.flatten()[:8].tolist()}
    # learning-rate schedule of utils.get_net (:265-273): HF get_scheduler("cosine_with_restarts", 1024 warm-up steps)
    from transformers import get_scheduler

    p = torch.nn.Parameter(torch.zeros(1))
    opt = ref_import.import_reference_module("arm.optim.lamb").Lamb([p], lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-5, adam=False)
    sched = get_scheduler("cosine_with_restarts", optimizer=opt, num_warmup_steps=4, num_training_steps=20)
    lrs = []
    for _ in range(20):
        p.grad = torch.ones(1)
        opt.step()
        sched.step()
        lrs.append(opt.param_groups[0]["lr"])
    out["lr_schedule_cosine_with_restarts_w4_t20"] = lrs
    path = os.path.join(ROOT, "tests", "golden", "train_golden.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, "ovssc frame columns:", out["ovssc_bal0"]["frame"]["columns"])
    print("ovssc stats", out["ovssc_bal0"]["stats"])
    print("vool stats", out["vool_bal1"]["stats"])


if __name__ == "__main__":
    main()
