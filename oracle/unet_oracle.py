"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32 functional ops) of the reference's 3-D UNet and SemAbs3D
forward.  Never imported by the product path.  Pinned by oracle/gen_golden_3d.py, which runs the UNMODIFIED
reference modules (unet3d.ResidualUNet3D, net.SemAbs3D / SemAbsVOOL from /root/reference) on the same seeded state
dicts and asserts agreement; the reference outputs are committed as tests/golden/unet_golden.npz.
State dicts use the reference's key names.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def _groups(channels, num_groups):
    # "use only one group if the given number of groups is greater than the number of channels" (unet3d.py:72-73)
    return 1 if channels < num_groups else num_groups


def _single_conv(sd, prefix, x, num_groups, relu):
    """SingleConv with order 'gc[r]' (create_conv, unet3d.py:20-95): GroupNorm -> Conv3d(3, pad 1, no bias) -> ReLU."""
    x = F.group_norm(x, _groups(x.shape[1], num_groups), sd[prefix + "groupnorm.weight"], sd[prefix + "groupnorm.bias"], 1e-5)
    x = F.conv3d(x, sd[prefix + "conv.weight"], None, padding=1)
    return F.relu(x) if relu else x


def _res_block(sd, prefix, x, num_groups):
    """ExtResNetBlock.forward (unet3d.py:243-259)."""
    o1 = _single_conv(sd, prefix + "conv1.", x, num_groups, True)
    o2 = _single_conv(sd, prefix + "conv2.", o1, num_groups, True)
    o3 = _single_conv(sd, prefix + "conv3.", o2, num_groups, False)
    return F.relu(o3 + o1)


def residual_unet3d(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_groups: int = 8, prefix: str = "") -> torch.Tensor:
    """Abstract3DUNet.forward for ResidualUNet3D (unet3d.py:596-621): MaxPool3d(2) between encoder blocks
    (:313-317), ConvTranspose3d(k3,s2,p1)(x, output_size=skip size) + skip sum in the decoder (:385-396,:438-440),
    final 1x1x1 conv (:578)."""
    n_enc = len({k[len(prefix + "encoders.") :].split(".")[0] for k in sd if k.startswith(prefix + "encoders.")})
    feats = []
    for i in range(n_enc):
        if i > 0:
            x = F.max_pool3d(x, 2)
        x = _res_block(sd, f"{prefix}encoders.{i}.basic_module.", x, num_groups)
        feats.insert(0, x)
    for j, skip in enumerate(feats[1:]):
        w, b = sd[f"{prefix}decoders.{j}.upsampling.upsample.weight"], sd[f"{prefix}decoders.{j}.upsampling.upsample.bias"]
        # output_size -> output_padding = target - ((in-1)*2 - 2 + 3)
        op = [skip.shape[2 + a] - ((x.shape[2 + a] - 1) * 2 - 2 + 3) for a in range(3)]
        x = F.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=op)
        x = _res_block(sd, f"{prefix}decoders.{j}.basic_module.", skip + x, num_groups)
    return F.conv3d(x, sd[prefix + "final_conv.weight"], sd[prefix + "final_conv.bias"])


# ---------------------------------------------------------------------------------------------------------
# SemAbs3D pieces (net.py)
# ---------------------------------------------------------------------------------------------------------
def point_grid_indices(points, bounds, grid_shape, cast_to_int=True):
    """VirtualGrid.get_points_grid_idxs (net.py:84-113): (p - lc) * (shape-1)/(uc-lc), trunc to int64, clamp."""
    lc = torch.tensor(bounds[0], dtype=torch.float32)
    uc = torch.tensor(bounds[1], dtype=torch.float32)
    scales = (torch.tensor(grid_shape, dtype=torch.float32) - 1) / (uc - lc)
    idx = (points + (-lc)) * scales
    if cast_to_int:
        idx = idx.to(torch.int64)
    out = torch.empty_like(idx)
    for i in range(3):
        out[..., i] = torch.clamp(idx[..., i], min=0, max=grid_shape[i] - 1)
    return out


def scatter_points_mean(xyz, feats, bounds, grid_shape):
    """VirtualGrid.scatter_points (net.py:185-201) with the reduce method the reference ends up using: MEAN
    (SemAbs3D builds its VirtualGrid without reduce_method and scatter_points ignores its argument);
    empty voxels are 0. Returns [B, C, X, Y, Z]."""
    B, npts, C = feats.shape
    idx = point_grid_indices(xyz, bounds, grid_shape)
    flat = (idx[..., 0] * grid_shape[1] + idx[..., 1]) * grid_shape[2] + idx[..., 2]
    n_vox = int(np.prod(grid_shape))
    vol = torch.zeros(B, n_vox, C).scatter_reduce(1, flat[..., None].expand(-1, -1, C), feats, reduce="mean", include_self=False)
    return vol.view(B, *grid_shape, C).permute(0, 4, 1, 2, 3).contiguous()


def point_mlp(sd, prefix, x):
    """pts_feat_extractor (net.py:358-367): Linear -> LeakyReLU(0.01) -> Linear -> LeakyReLU -> Linear."""
    x = F.leaky_relu(F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    x = F.leaky_relu(F.linear(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"]))
    return F.linear(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"])


def implicit_decoder(sd, prefix, vol, query, bounds, grid_shape, concat_xyz=False):
    """ImplicitVolumetricDecoder.forward (net.py:215-256), quirks included: indices are divided by grid_shape (not
    shape-1) before mapping to [-1,1], and (x,y,z) is handed to grid_sample whose last-dim convention is (W,H,D),
    so the [N,C,X,Y,Z] volume is read with x and z swapped."""
    q = point_grid_indices(query, bounds, grid_shape, cast_to_int=False).float()
    for i in range(3):
        q[..., i] = q[..., i] / grid_shape[i]
    qn = 2.0 * q - 1.0
    grid = qn.view(*qn.shape[:2], 1, 1, 3)
    s = F.grid_sample(vol, grid, mode="bilinear", padding_mode="border", align_corners=True)
    s = s.view(s.shape[:3]).permute(0, 2, 1).contiguous()
    Bn, L, C = s.shape
    s = s.view(Bn * L, C)
    if concat_xyz:
        s = torch.cat((s, grid.view(Bn * L, 3)), dim=-1)
    h = F.leaky_relu(F.linear(s, sd[prefix + "mlp.0.weight"], sd[prefix + "mlp.0.bias"]))
    return F.linear(h, sd[prefix + "mlp.2.weight"], sd[prefix + "mlp.2.bias"]).view(Bn, L, -1)


def semabs3d_forward(sd, input_xyz_pts, input_feature_pts, output_xyz_pts, bounds, grid_shape, num_groups=8,
                     concat_xyz=False, prefix="", return_volume=False):
    """SemAbs3D.forward (net.py:383-439), default network_inputs=["saliency"], use_pts_feat_extractor=True."""
    B, P, npts = input_feature_pts.shape[:3]
    xyz = input_xyz_pts.unsqueeze(1).repeat(1, P, 1, 1).view(B * P, npts, 3)
    feat = input_feature_pts.view(B * P, npts, -1)
    feat = point_mlp(sd, prefix + "pts_feat_extractor.", torch.cat((xyz, feat), dim=-1))
    vol = scatter_points_mean(xyz, feat, bounds, grid_shape)
    vol = residual_unet3d(sd, vol, num_groups, prefix=prefix + "vol_feature_extractor.")
    nq = output_xyz_pts.shape[2]
    out = implicit_decoder(sd, prefix + "visual_sampler.", vol, output_xyz_pts.view(B * P, nq, 3), bounds, grid_shape, concat_xyz)
    out = out.view(B, P, nq, -1).squeeze(-1)
    return (out, vol) if return_volume else out


RELATIONS = ["in", "behind", "in front of", "on the left of", "on the right of", "on", "[pad]"]


def semabsvool_forward(sd, input_xyz_pts, target_saliency, reference_saliency, output_xyz_pts,
                       spatial_relation_name: Sequence[Sequence[str]], bounds, grid_shape, num_groups=8,
                       concat_xyz=False):
    """SemAbsVOOL.forward (net.py:528-579) with pointing_method='cosine_sim' (PointingAttention.cosine_sim,
    net.py:300-309, temperature 0.07)."""
    B, num_descs = np.array(spatial_relation_name).T.shape
    place = torch.zeros_like(input_xyz_pts)[..., None, 0:1, :].repeat(1, num_descs, 1, 1)
    vols = []
    for sal in (target_saliency, reference_saliency):
        # the completion net is built without decoder_concat_xyz_pts (net.py:481: SemAbs3D(device=device, **kwargs))
        _, v = semabs3d_forward(sd, input_xyz_pts, sal, place, bounds, grid_shape, num_groups, False,
                                prefix="completion_net.", return_volume=True)
        vols.append(v)
    fv = torch.cat(vols, dim=1)
    nq = output_xyz_pts.shape[-2]
    key = implicit_decoder(sd, "spatial_sampler.", fv, output_xyz_pts.view(B * num_descs, nq, 3), bounds, grid_shape, concat_xyz)
    emb = torch.stack([torch.stack([sd["relation_embeddings." + spatial_relation_name[d][b]] for b in range(B)], 0)
                       for d in range(num_descs)], 0).permute(1, 0, 2).contiguous()  # fmt: skip
    query = emb.view(B * num_descs, 1, -1)
    return (torch.cosine_similarity(key, query, dim=-1) / 0.07).view(B, num_descs, nq)
