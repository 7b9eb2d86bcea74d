"""Evaluation metrics on the device — mirrors of `utils.prediction_analysis` (utils.py:338-380) and
`utils.voxelize_points` (utils.py:617-665), the two pieces `train_ovssc.get_detailed_stats` (train_ovssc.py:20-78) spends
its time in (a Python double loop over (scene, patch) with a dozen tiny torch ops and `.item()` syncs each, and three
torch_scatter calls).  Here: one counting kernel per call; the per-(scene, patch) ratios are formed from 7 integers."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import ops
from ._lib import check, i32, lib, ptr, stream_ptr
from .net import VirtualGrid


def _u8(t: torch.Tensor, dev) -> torch.Tensor:
    return t.to(dev).bool().to(torch.uint8).contiguous()


def _ratios(counts: torch.Tensor) -> Dict[str, List[float]]:
    """counts int64 [N,7] -> the reference's lists (NaN exactly where the reference yields NaN: precision / recall for
    an empty denominator via its explicit np.NAN, iou and the two means via 0/0)."""
    c = counts.cpu().double()
    tp, pp, lp, un, fn, fp, kept = (c[:, i] for i in range(7))
    nan = torch.full_like(tp, float("nan"))
    div = lambda a, b: torch.where(b != 0, a / b.clamp_min(1), nan)
    return {"precision": div(tp, pp).tolist(), "recall": div(tp, lp).tolist(), "false_negative": div(fn, kept).tolist(),
            "false_positive": div(fp, kept).tolist(), "iou": div(tp, un).tolist()}


def confusion_counts(prediction, label, ignore, device="cuda") -> torch.Tensor:
    B, P, n = prediction.shape
    counts = torch.empty(B * P, 7, dtype=torch.int64, device=device)
    p8, l8, g8 = _u8(prediction, device), _u8(label, device), _u8(ignore, device)  # (kept alive until the launch is queued)
    check(lib().semabs_confusion_counts(ptr(p8), ptr(l8), ptr(g8), i32(B * P), C.c_int64(n), ptr(counts), stream_ptr()))
    return counts


def prediction_analysis(prediction, label, ignore, device="cuda") -> Dict[str, List[float]]:
    """Same signature / return value as utils.prediction_analysis: dict of python lists, (scene, patch) row-major."""
    return _ratios(confusion_counts(prediction, label, ignore, device))


def _voxelized(prediction, label, xyz_pts, voxel_shape, scene_bounds, ignore_pts, device, want_volumes):
    B, P, n = prediction.shape
    N = B * P
    vg = VirtualGrid(scene_bounds=np.asarray(scene_bounds), grid_shape=voxel_shape, batch_size=N)
    neg_lc, scale, shape = vg.kernel_grid()
    V = int(np.prod(voxel_shape))
    xyz = xyz_pts.to(device).float().reshape(N, n, 3).contiguous()
    flags = torch.empty(N, V, dtype=torch.int32, device=device)
    counts = torch.empty(N, 7, dtype=torch.int64, device=device)
    vols = [torch.empty(N, V, dtype=torch.uint8, device=device) for _ in range(3)] if want_volumes else [None] * 3
    p8, l8, g8 = _u8(prediction, device), _u8(label, device), _u8(ignore_pts, device)
    check(lib().semabs_voxelized_confusion_counts(ptr(xyz), i32(1), ptr(p8), ptr(l8), ptr(g8), i32(N), C.c_int64(n),
                                                  ops._host3(neg_lc, ops._cf),
                                                  ops._host3(scale, ops._cf), ops._host3(shape, ops._ci), ptr(flags), ptr(counts),
                                                  ptr(vols[0]), ptr(vols[1]), ptr(vols[2]), stream_ptr()))
    return counts, vols


def voxelize_points(prediction, label, xyz_pts, voxel_shape: Tuple[int, int, int], scene_bounds, ignore_pts, device="cuda"):
    """utils.voxelize_points: {"prediction" bool, "label" float, "ignore" bool}, each [batch, patches, prod(voxel_shape)]."""
    B, P, _ = prediction.shape
    _, (vp, vl, vi) = _voxelized(prediction, label, xyz_pts, voxel_shape, scene_bounds, ignore_pts, device, True)
    return {"prediction": vp.bool().view(B, P, -1), "label": vl.float().view(B, P, -1), "ignore": vi.bool().view(B, P, -1)}


def voxel_prediction_analysis(prediction, label, xyz_pts, voxel_shape, scene_bounds, ignore_pts, device="cuda"):
    """prediction_analysis(**voxelize_points(...)) (train_ovssc.py:44-62) without materialising the voxel volumes."""
    counts, _ = _voxelized(prediction, label, xyz_pts, voxel_shape, scene_bounds, ignore_pts, device, False)
    return _ratios(counts)
