"""RGB-D -> relevancy -> OVSSC voxel logits, device resident (BASELINE.json configs[4]).

Mirror of the inference glue the reference spreads over visualize.py: `prep_data` (:60-152: relevancy maps x50, minus the
mean over labels, back-projected depth, in-bounds mask, per-class point features) and `process_batch_ovssc` (:157-211:
sub-sample `num_input_pts` points per class, query lattice from `get_sample_points` :283-298, SemAbs3D logits) with
the class arg-max of :236-238.  Differences, on purpose: the relevancy maps never leave the GPU; the UNet runs ONCE per
image for all classes (the reference re-runs it for every 2^20-point chunk of the lattice, :182-211) and only the implicit
decoder is chunked; the point sub-sample is drawn once per class instead of once per chunk.  The cutoff / frustum / TSDF
masking of :212-247 is `prediction_volumes` below (bit-identical to the reference's fusion.TSDFVolume + check_pts_in_frustum,
tests/test_pipeline.py).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .clip import ClipWrapper
from .net import SemAbs3D

PROMPT = "a photograph of a {} in a home."


def back_project(depth: torch.Tensor, cam_intr, cam_pose=None) -> torch.Tensor:
    """point_cloud.get_pointcloud (point_cloud.py:34-67) on the device: depth [H,W] fp32 -> xyz [H*W,3] in camera (or
    world, with a 3x4 / 4x4 `cam_pose`) coordinates, pixel order row-major like the reference's reshape(-1,3)."""
    dev = depth.device
    H, W = depth.shape
    K = torch.as_tensor(np.asarray(cam_intr), dtype=torch.float32, device=dev)
    px = torch.linspace(0, W - 1, W, device=dev).view(1, W).expand(H, W)
    py = torch.linspace(0, H - 1, H, device=dev).view(H, 1).expand(H, W)
    x = (px - K[0, 2]) * (depth / K[0, 0])
    y = (py - K[1, 2]) * (depth / K[1, 1])
    pts = torch.stack([x, y, depth], dim=-1).reshape(-1, 3)
    if cam_pose is not None:
        T = torch.as_tensor(np.asarray(cam_pose), dtype=torch.float32, device=dev)
        pts = pts @ T[:3, :3].T + T[:3, 3]
    return pts


def filter_pts_bounds(xyz: torch.Tensor, bounds) -> torch.Tensor:
    """point_cloud.filter_pts_bounds (point_cloud.py:24-31)."""
    lo = torch.as_tensor(np.asarray(bounds[0]), dtype=xyz.dtype, device=xyz.device)
    hi = torch.as_tensor(np.asarray(bounds[1]), dtype=xyz.dtype, device=xyz.device)
    return ((xyz >= lo) & (xyz <= hi)).all(dim=-1)


def get_sample_points(sampling_shape: Sequence[int], scene_bounds, device) -> torch.Tensor:
    """visualize.get_sample_points (visualize.py:283-298): lattice of query points, [prod(shape), 3]."""
    axes = [torch.arange(0, n, device=device) for n in sampling_shape]
    idx = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).to(torch.float32)
    lc = torch.tensor(scene_bounds[0], device=device, dtype=torch.float32)
    uc = torch.tensor(scene_bounds[1], device=device, dtype=torch.float32)
    scales = (uc - lc) / (torch.tensor(list(sampling_shape), device=device, dtype=torch.float32) - 1)
    return (idx * scales + lc).view(-1, 3)


@torch.no_grad()
def relevancy_point_features(rgb: np.ndarray, labels: Sequence[str], saliency_config: dict, subtract_mean: bool = True,
                             prompts=(PROMPT,)) -> torch.Tensor:
    """prep_data's relevancy half (visualize.py:93-110): maps [P,H,W] on the device, x50, minus the mean over labels."""
    ClipWrapper.check_initialized()
    gc = ClipWrapper.clip_gradcam
    gc.templates = list(prompts)
    gc.set_classes(list(labels))
    rel = ClipWrapper.get_clip_saliency_convolve(img=rgb, text_labels=list(labels), keep_on_device=True, **saliency_config) * 50
    if subtract_mean:
        rel = rel - rel.mean(dim=0, keepdim=True)
    return rel


@torch.no_grad()
def rgbd_to_ovssc_logits(net: SemAbs3D, rgb: np.ndarray, depth, cam_intr, cam_extr, labels: Sequence[str],
                         scene_bounds, saliency_config: dict, sampling_shape: Tuple[int, int, int] = (240, 240, 240),
                         num_input_pts: int = 80000, num_pts_per_pass: int = 2**20, subtract_mean: bool = True,
                         generator: Optional[torch.Generator] = None, masked_volumes: bool = False,
                         cutoff: float = -3.0) -> Dict[str, torch.Tensor]:
    """One image through the whole path. Returns {"relevancies" [P,H,W], "logits" [P, *sampling_shape],
    "prediction" int64 [*sampling_shape] (arg-max class, visualize.py:236)} — all on the device.  Defaults are the
    reference's (visualize.py:163-164: a 240^3 lattice = 13.8 M query points per class in 2^20-point chunks)."""
    dev = next(net.parameters()).device
    P = len(labels)
    rel = relevancy_point_features(rgb, labels, saliency_config, subtract_mean)
    depth_t = torch.as_tensor(depth, dtype=torch.float32, device=dev)
    xyz = back_project(depth_t, cam_intr, cam_extr)
    idx_in = filter_pts_bounds(xyz, scene_bounds).nonzero().squeeze(1)
    if idx_in.numel() == 0:
        raise ValueError("no depth point falls inside scene_bounds")
    # np.random.choice(n, size=num_input_pts) (with replacement) per class, visualize.py:189
    pick = torch.randint(0, idx_in.numel(), (P, num_input_pts), device=dev, generator=generator)
    sel = idx_in[pick]                                   # [P, npts] pixel indices
    feats = torch.gather(rel.view(P, -1), 1, sel)        # [P, npts]
    grid_points = get_sample_points(sampling_shape, scene_bounds, dev)
    nq = grid_points.shape[0]
    logits = torch.empty(P, nq, device=dev)
    # every class has its own point sub-sample, so xyz is a per-class batch (B = P scenes of one patch each); the UNet
    # runs once, only the implicit decoder walks the lattice in chunks
    vol = net.feature_volume(xyz[sel].contiguous(), feats.view(P, 1, num_input_pts, 1).contiguous())
    C = net.unet_num_channels
    for j in range(0, nq, num_pts_per_pass):
        q = grid_points[j : j + num_pts_per_pass]
        out = net.visual_sampler.run([vol], C, net.vg, q.unsqueeze(0).expand(P, -1, -1).contiguous())
        logits[:, j : j + q.shape[0]] = out.view(P, -1)
    logits = logits.view(P, *sampling_shape)
    out = {"relevancies": rel, "logits": logits, "prediction": logits.argmax(dim=0)}
    if masked_volumes:  # visualize.py:212-247: cutoff / frustum / TSDF masking of the per-class volumes
        out["prediction_volumes"] = prediction_volumes(logits, sampling_shape, scene_bounds, depth_t, cam_intr, cam_extr, cutoff)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# post-processing of the dense sweep (visualize.process_batch_ovssc, visualize.py:212-247): empty / frustum / TSDF masks
# ---------------------------------------------------------------------------------------------------------------------
def check_pts_in_frustum(xyz_pts: torch.Tensor, depth_shape, cam_pose, cam_intr) -> torch.Tensor:
    """point_cloud.check_pts_in_frustum (point_cloud.py:87-110): world points -> camera frame (inverse pose, fp64 like the
    numpy original) -> pinhole projection; inside iff 0 <= x < w, 0 <= y < h and z > 0."""
    dev = xyz_pts.device
    T = torch.linalg.inv(torch.as_tensor(np.asarray(cam_pose), dtype=torch.float64, device=dev))
    K = torch.as_tensor(np.asarray(cam_intr), dtype=torch.float64, device=dev)
    p = xyz_pts.to(torch.float64) @ T[:3, :3].T + T[:3, 3]
    z = p[:, 2]
    px = (K[0, 0] / z) * p[:, 0] + K[0, 2]
    py = (K[1, 1] / z) * p[:, 1] + K[1, 2]
    h, w = depth_shape
    return (px >= 0) & (px < w) & (py >= 0) & (py < h) & (z > 0)


def tsdf_single_frame(scene_bounds, voxel_size: float, depth: torch.Tensor, cam_intr, cam_pose) -> torch.Tensor:
    """fusion.TSDFVolume(vol_bnds, voxel_size).integrate(one frame).get_volume()[0] (fusion.py:10-173): the truncated signed
    distance of every voxel to the observed surface along the camera ray, -1 where nothing was observed.  Same dtypes as
    the original: voxel centres in fp32, camera transform in fp64, fp32 intrinsics, round-half-even pixel lookup."""
    dev = depth.device
    b = np.asarray(scene_bounds, dtype=np.float64).T.copy()                    # (3, 2) like vol_bnds
    vol_dim = np.ceil((b[:, 1] - b[:, 0]) / float(voxel_size)).astype(int)
    origin = torch.as_tensor(b[:, 0].astype(np.float32), device=dev)
    trunc = 5 * float(voxel_size)
    ijk = torch.stack(torch.meshgrid(*[torch.arange(int(n), device=dev) for n in vol_dim], indexing="ij"), dim=-1).reshape(-1, 3)
    world = (origin.double() + float(voxel_size) * ijk.double()).float()       # vox2world writes fp32
    T = torch.linalg.inv(torch.as_tensor(np.asarray(cam_pose), dtype=torch.float64, device=dev))
    cam = world.double() @ T[:3, :3].T + T[:3, 3]
    K = torch.as_tensor(np.asarray(cam_intr), dtype=torch.float32, device=dev).double()  # cam2pix casts intrinsics to fp32
    z = cam[:, 2]
    px = torch.round(cam[:, 0] * K[0, 0] / z + K[0, 2])
    py = torch.round(cam[:, 1] * K[1, 1] / z + K[1, 2])
    h, w = depth.shape
    valid_pix = (px >= 0) & (px < w) & (py >= 0) & (py < h) & (z > 0)
    pxi = torch.where(valid_pix, px, torch.zeros_like(px)).long()
    pyi = torch.where(valid_pix, py, torch.zeros_like(py)).long()
    depth_val = torch.where(valid_pix, depth.double()[pyi, pxi], torch.zeros_like(z))
    diff = depth_val - z
    valid = (depth_val > 0) & (diff >= -trunc)
    dist = torch.clamp(diff / trunc, -1.0, 1.0)
    tsdf = torch.where(valid, dist, torch.full_like(dist, -1.0)).float()
    return tsdf.view(*[int(n) for n in vol_dim])


@torch.no_grad()
def prediction_volumes(logits: torch.Tensor, sampling_shape, scene_bounds, depth, cam_intr, cam_extr, cutoff: float = -3.0):
    """The tail of visualize.process_batch_ovssc (visualize.py:212-247): arg-max class per lattice point, zeroed where every
    class is below `cutoff`, outside the camera frustum, or in observed free space (TSDF > 0).  logits [P, *sampling_shape]
    -> float volumes [P, *sampling_shape] (1 = the class is predicted there)."""
    dev = logits.device
    P = logits.shape[0]
    depth_t = torch.as_tensor(depth, dtype=torch.float32, device=dev)
    grid_points = get_sample_points(sampling_shape, scene_bounds, dev)
    voxel_size = (scene_bounds[1][0] - scene_bounds[0][0]) / sampling_shape[0]
    tsdf = tsdf_single_frame(scene_bounds, voxel_size, depth_t, cam_intr, cam_extr)
    assert tuple(tsdf.shape) == tuple(sampling_shape), "the TSDF grid of visualize.py:212-216 must coincide with the lattice"
    logprobs = logits.permute(*range(1, logits.dim()), 0)                      # [..., P]
    prediction = logprobs.argmax(dim=-1)
    empty = (logprobs < cutoff).all(dim=-1)
    in_frustum = check_pts_in_frustum(grid_points, depth_t.shape, cam_extr, cam_intr).view(*sampling_shape)
    keep = (~empty) & in_frustum & ~(tsdf > 0.0)
    return torch.stack([((prediction == c) & keep).float() for c in range(P)])


@torch.no_grad()
def process_batch_ovssc(net: SemAbs3D, batch: Dict, scene_bounds, device, num_input_pts: int,
                        sampling_shape: Tuple[int, int, int] = (240, 240, 240), num_pts_per_pass: int = int(2**20),
                        cutoff: float = -3.0, generator: Optional[torch.Generator] = None, return_logits: bool = False):
    """visualize.process_batch_ovssc (visualize.py:157-248) with the reference's signature and batch keys
    (`input_xyz_pts` [1, n, 3], `input_feature_pts` [1, P, n, 1] or [P, n], `ovssc_obj_classes`, `depth`, `cam_intr`,
    `cam_extr`): per class a random sub-sample of `num_input_pts` input points (np.random.choice with replacement there; a
    device draw here), logits on the `sampling_shape` lattice, then the arg-max / cutoff / frustum / TSDF masks.  Returns
    {class label: float32 numpy volume [*sampling_shape]} like the reference.  The UNet runs once per class batch (the
    reference re-runs it, with a fresh sub-sample, for each of the 14 lattice chunks); only the implicit decoder is chunked."""
    dev = torch.device(device)
    classes = list(batch["ovssc_obj_classes"])
    P = len(classes)
    xyz_all = torch.as_tensor(batch["input_xyz_pts"], dtype=torch.float32, device=dev).reshape(-1, 3)
    feats_all = torch.as_tensor(batch["input_feature_pts"], dtype=torch.float32, device=dev).reshape(P, -1)
    n = xyz_all.shape[0]
    pick = torch.randint(0, n, (P, num_input_pts), device=dev, generator=generator)
    feats = torch.gather(feats_all, 1, pick)
    vol = net.feature_volume(xyz_all[pick].contiguous(), feats.view(P, 1, num_input_pts, 1).contiguous())
    grid_points = get_sample_points(sampling_shape, scene_bounds, dev)
    assert bool(filter_pts_bounds(grid_points, scene_bounds).all())
    nq = grid_points.shape[0]
    logits = torch.empty(P, nq, device=dev)
    for j in range(0, nq, num_pts_per_pass):
        q = grid_points[j : j + num_pts_per_pass]
        out = net.visual_sampler.run([vol], net.unet_num_channels, net.vg, q.unsqueeze(0).expand(P, -1, -1).contiguous())
        logits[:, j : j + q.shape[0]] = out.view(P, -1)
    vols = prediction_volumes(logits.view(P, *sampling_shape), sampling_shape, scene_bounds, batch["depth"], batch["cam_intr"],
                              batch["cam_extr"], cutoff)
    out = {label: vols[c].cpu().numpy() for c, label in enumerate(classes)}
    return (out, logits.view(P, *sampling_shape)) if return_logits else out
