"""`train_ovssc` entry points on the B200 kernels — mirror of the reference's train_ovssc.py: `get_detailed_stats`
(train_ovssc.py:11-78), `get_losses(net, batch, cutoffs=[0], balance_positive_negative=False, **kw) -> (stats, DataFrame)`
(:81-169), the `approach` table (:172-175) and the `python train_ovssc.py --file_path ... --log ...` command (:178-197).

The loss / accuracy are one fused kernel (`semabs_bce_with_logits`: masked BCE-with-logits, its gradient and the accuracy
in a single pass); every row of the per-cutoff IoU table comes from the counting kernels of `semabs_b200.metrics`
(identical numbers to utils.prediction_analysis / utils.voxelize_points, NaN cases included)."""
from __future__ import annotations

from typing import Dict, Tuple, Union

import numpy as np
import pandas as pd
import torch

from . import metrics
from .net import SemAbs3D
from .train import _BceFn, get_bce_weight

_METRICS = ("precision", "recall", "false_negative", "false_positive", "iou")


def _analysis_columns(prefix: str, table: Dict[str, list]) -> Dict[str, list]:
    return {f"{prefix}{k}": table[k] for k in _METRICS}


def get_detailed_stats(prediction, gt_label, xyz_pts, patch_labels, scene_ids, scene_bounds, ignore_pts,
                       detailed_analysis=False, eval_device="cuda", **kwargs) -> pd.DataFrame:
    """One row per (scene, patch): point-wise and 32^3-voxelised (optionally 64^3) precision / recall / false-negative /
    false-positive / IoU; padding patches (label "") get NaN in every metric column."""
    num_scenes, num_patches = patch_labels.shape
    cols = {"scene_id": [s for s in scene_ids for _ in range(num_patches)], "label": patch_labels.reshape(-1).tolist()}
    cols.update(_analysis_columns("point_", metrics.prediction_analysis(prediction, gt_label, ignore_pts, device=eval_device)))
    for res in (32, 64) if detailed_analysis else (32,):
        cols.update(_analysis_columns(f"voxel{res}x{res}x{res}_", metrics.voxel_prediction_analysis(
            prediction, gt_label, xyz_pts, (res, res, res), scene_bounds, ignore_pts, device=eval_device)))
    pad = [i for i, label in enumerate(cols["label"]) if label == ""]
    for k, v in cols.items():
        if "voxel" in k or "point" in k:
            for i in pad:
                v[i] = np.nan
    return pd.DataFrame.from_dict(cols)


def scene_bounds_of(net):
    """the network's own voxel-grid bounds: default for `scene_bounds` when the caller (unlike utils.train, which forwards
    every parsed flag) does not pass it"""
    vg = getattr(net, "vg", None) or getattr(getattr(net, "completion_net", None), "vg", None)
    if vg is None:
        raise TypeError("get_losses: pass scene_bounds=... (the network exposes no voxel grid to take it from)")
    return [list(vg.lower_corner), list(vg.upper_corner)]


def _forward_in_patch_chunks(net, batch):
    """The reference's fallback for > 500 000 query points (train_ovssc.py:93-126): one patch per forward."""
    P = batch["output_xyz_pts"].shape[1]
    outs = []
    for i in range(P):
        sub = dict(batch, output_xyz_pts=batch["output_xyz_pts"][:, i : i + 1])
        if batch["input_feature_pts"].shape[1] == P:
            sub["input_feature_pts"] = batch["input_feature_pts"][:, i : i + 1]
        if "semantic_class_features" in batch:
            sub["semantic_class_features"] = batch["semantic_class_features"][:, i : i + 1]
        outs.append(net(**sub))
    return torch.cat(outs, dim=1)


def get_losses(net, batch: dict, cutoffs=[0], balance_positive_negative: bool = False,
               **kwargs) -> Tuple[Dict[str, Union[float, torch.Tensor]], pd.DataFrame]:
    stats = {}
    if "scene_bounds" not in kwargs:
        kwargs["scene_bounds"] = scene_bounds_of(net)
    outputs = net(**batch) if batch["output_xyz_pts"].shape[2] <= 500000 else _forward_in_patch_chunks(net, batch)
    # like the reference, the collated [P][B] label lists and the out-of-bounds mask are normalised IN the batch dict
    if not isinstance(batch["patch_labels"], np.ndarray):  # (idempotent: a batch dict that is reused arrives normalised)
        batch["patch_labels"] = np.array(batch["patch_labels"]).T
    batch["out_of_bounds_pts"] = batch["out_of_bounds_pts"].view(outputs.shape)
    ignore = torch.zeros_like(outputs, dtype=torch.bool)
    ignore[torch.from_numpy(batch["patch_labels"] == "").to(outputs.device)] = True          # padding patches
    ignore |= batch["out_of_bounds_pts"].bool() | batch["out_of_frustum_pts_mask"].view(outputs.shape).bool()
    labels = batch["output_label_pts"]
    weight = get_bce_weight(labels, balance_positive_negative)
    stats["loss"], stats["accuracy"] = _BceFn.apply(outputs.contiguous(), labels, weight, ignore)
    with torch.no_grad():
        frames = []
        for cutoff in cutoffs:
            df = get_detailed_stats(prediction=outputs > cutoff, gt_label=labels.bool(), xyz_pts=batch["output_xyz_pts"],
                                    ignore_pts=ignore, patch_labels=batch["patch_labels"], scene_ids=batch["scene_id"],
                                    eval_device=getattr(net, "device", outputs.device), **kwargs)
            df["cutoff"] = [cutoff] * len(df)
            frames.append(df)
        detailed_stats = pd.concat(frames)
        for k in detailed_stats.columns:
            if "iou" in k:
                stats[k] = detailed_stats[k].mean()
    return stats, detailed_stats


approach = {"semantic_abstraction": SemAbs3D}


def main(argv=None):
    from . import utils

    parser = utils.config_parser()
    parser.add_argument("--log", type=str, required=True)
    parser.add_argument("--approach", choices=approach.keys(), default="semantic_abstraction")
    parser.add_argument("--synthetic_scenes", type=int, default=8)
    args = parser.parse_args(argv)
    exp = utils.setup_experiment(args=args, net_class=approach[args.approach], dataset_class=utils.SyntheticOVSSCDataset,
                                 length=args.synthetic_scenes)
    return utils.train(get_losses_fn=get_losses, **exp, **vars(args))


if __name__ == "__main__":
    main()
