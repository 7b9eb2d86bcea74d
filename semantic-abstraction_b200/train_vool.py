"""`train_vool` entry points on the B200 kernels — mirror of the reference's train_vool.py: `get_detailed_stats`
(train_vool.py:15-115), `get_losses(net, batch, cutoffs=[-2.0], balance_positive_negative=False, **kw) ->
(stats, DataFrame)` (:118-206), the `approach` table (:209-213) and the `python train_vool.py ...` command (:215-233).

Unlike train_ovssc the LOSS runs over every point (train_vool.py:172-179); only the accuracy and the IoU table honour the
padding ("[pad]" relation) / out-of-bounds mask."""
from __future__ import annotations

from typing import Dict, Tuple, Union

import numpy as np
import pandas as pd
import torch

from . import metrics
from .net import SemAbsVOOL
from .train import _BceFn, bce_with_logits_masked, get_bce_weight
from .train_ovssc import _analysis_columns, scene_bounds_of


def get_detailed_stats(prediction, gt_label, xyz_pts, scene_ids, target_obj_names, reference_obj_names, spatial_relation_names,
                       scene_bounds, ignore_pts, detailed_analysis=False, eval_device="cuda", **kwargs) -> pd.DataFrame:
    """One row per (scene, description). The reference voxelises 10 descriptions at a time to bound memory
    (train_vool.py:52-71); the counting kernel needs no volumes, so all descriptions go in one launch (same numbers)."""
    num_scenes, num_descs = gt_label.shape[:2]
    flat = lambda names: np.array(names).T.reshape(-1).tolist()
    cols = {"scene_id": [s for s in scene_ids for _ in range(num_descs)], "target_obj_name": flat(target_obj_names),
            "reference_obj_name": flat(reference_obj_names), "spatial_relation_name": flat(spatial_relation_names)}
    cols.update(_analysis_columns("point_", metrics.prediction_analysis(prediction, gt_label, ignore_pts, device=eval_device)))
    for res in (32, 64) if detailed_analysis else (32,):
        cols.update(_analysis_columns(f"voxel{res}x{res}x{res}_", metrics.voxel_prediction_analysis(
            prediction, gt_label, xyz_pts, (res, res, res), scene_bounds, ignore_pts, device=eval_device)))
    pad = [i for i, rel in enumerate(cols["spatial_relation_name"]) if rel == "[pad]"]
    for k, v in cols.items():
        if "voxel" in k or "point" in k:
            for i in pad:
                v[i] = np.nan
    return pd.DataFrame.from_dict(cols)


def _forward_in_description_chunks(net, batch):
    """> 500 000 query points: one description per forward (train_vool.py:126-160)."""
    D = batch["output_label_pts"].shape[1]
    names = np.array(batch["spatial_relation_name"]).T  # [B, D]
    outs = []
    for i in range(D):
        sub = dict(batch, spatial_relation_name=names[:, i : i + 1].T)
        for k in ("input_target_saliency_pts", "input_reference_saliency_pts", "input_description_saliency_pts", "output_xyz_pts"):
            if k in batch:
                sub[k] = batch[k][:, i : i + 1]
        outs.append(net(**sub))
    return torch.cat(outs, dim=1)


def get_losses(net, batch: dict, cutoffs=[-2.0], balance_positive_negative: bool = False,
               **kwargs) -> Tuple[Dict[str, Union[float, torch.Tensor]], pd.DataFrame]:
    stats = {}
    if "scene_bounds" not in kwargs:
        kwargs["scene_bounds"] = scene_bounds_of(net)
    labels = batch["output_label_pts"]
    outputs = net(**batch) if labels.shape[2] <= 500000 else _forward_in_description_chunks(net, batch)
    ignore = torch.zeros_like(outputs, dtype=torch.bool)
    ignore[torch.from_numpy(np.array(batch["spatial_relation_name"]).T == "[pad]").to(outputs.device)] = True
    ignore |= batch["out_of_bounds_pts"].view(outputs.shape).bool()
    weight = get_bce_weight(labels, balance_positive_negative)
    stats["loss"], _ = _BceFn.apply(outputs.contiguous(), labels, weight, None)
    with torch.no_grad():
        _, stats["accuracy"], _ = bce_with_logits_masked(outputs.detach().contiguous(), labels, None, ignore, need_grad=False)
        frames = []
        for cutoff in cutoffs:
            df = get_detailed_stats(prediction=outputs > cutoff, gt_label=labels.bool(), xyz_pts=batch["output_xyz_pts"],
                                    ignore_pts=ignore, target_obj_names=batch["target_obj_name"],
                                    reference_obj_names=batch["reference_obj_name"],
                                    spatial_relation_names=batch["spatial_relation_name"], scene_ids=batch["scene_id"],
                                    eval_device=getattr(net, "device", outputs.device), **kwargs)
            df["cutoff"] = [cutoff] * len(df)
            frames.append(df)
        detailed_stats = pd.concat(frames)
        for k in detailed_stats.columns:
            if "iou" in k:
                stats[k] = detailed_stats[k].mean()
    return stats, detailed_stats


approach = {"semantic_abstraction": SemAbsVOOL}


def main(argv=None):
    from . import utils

    parser = utils.config_parser()
    parser.add_argument("--log", type=str, required=True)
    parser.add_argument("--approach", choices=approach.keys(), default="semantic_abstraction")
    parser.add_argument("--synthetic_scenes", type=int, default=8)
    args = parser.parse_args(argv)
    exp = utils.setup_experiment(args=args, net_class=approach[args.approach], dataset_class=utils.SyntheticVOOLDataset,
                                 length=args.synthetic_scenes)
    return utils.train(get_losses_fn=get_losses, **exp, **vars(args))


if __name__ == "__main__":
    main()
