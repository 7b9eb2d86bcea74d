"""SemAbs3D / SemAbsVOOL on the B200 kernels — host-side mirror of the reference classes of the same names
(reference net.py: VirtualGrid :23-201, ImplicitVolumetricDecoder :204-256, PointingAttention :259-316, SemAbs3D
:319-439, SemAbsVOOL :469-579).  Constructor arguments, forward() keyword arguments, the `steps` buffer and every
state-dict key are the reference's, so released `ovssc.pth` / `vool.pth` checkpoints load with load_state_dict.

The nn.Module containers own parameters only (built in the reference's order, so the same torch seed draws the
same initial values); forward() issues C-ABI kernels on channels-last device buffers:
  points -> [point MLP + scatter-MEAN voxelisation + GroupNorm stats] -> ResidualUNet3D (channels-last) ->
  [trilinear gather + decoder MLP (+ cosine-similarity pointing head)].
Reference quirks reproduced on purpose: voxelisation reduces with MEAN although the config
says max; the decoder normalises indices by `shape` (not shape-1) and samples the volume with x and z swapped.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
from torch.nn import LeakyReLU, Linear, Module, ParameterDict, Sequential
from torch.nn.parameter import Parameter

from . import ops
from .unet3d import ResidualUNet3D, _pad16

F32, F64 = torch.float32, torch.float64


class VirtualGrid:
    """Index arithmetic of the reference VirtualGrid (net.py:23-133) — host-side description only."""

    def __init__(self, scene_bounds, grid_shape=(32, 32, 32), batch_size=8, device=torch.device("cpu"),
                 int_dtype=torch.int64, float_dtype=torch.float32, reduce_method="mean"):
        self.lower_corner = tuple(float(v) for v in scene_bounds[0])
        self.upper_corner = tuple(float(v) for v in scene_bounds[1])
        self.grid_shape = tuple(int(v) for v in grid_shape)
        self.batch_size = int(batch_size)
        self.device = device
        self.reduce_method = reduce_method  # SemAbs3D never overrides it -> "mean" (net.py:339-344)

    def kernel_grid(self):
        """(-lc, (shape-1)/(uc-lc), shape) with the scale computed in fp32 exactly like net.py:92-98."""
        lc = np.asarray(self.lower_corner, dtype=np.float32)
        uc = np.asarray(self.upper_corner, dtype=np.float32)
        scale = (np.asarray(self.grid_shape, dtype=np.float32) - np.float32(1)) / (uc - lc)
        return [float(v) for v in -lc], [float(v) for v in scale], list(self.grid_shape)

    @property
    def num_grids(self):
        return int(np.prod((self.batch_size,) + self.grid_shape))


class ImplicitVolumetricDecoder(Module):
    def __init__(self, hidden_size: int, output_dim: int, concat_xyz_pts: bool = False):
        super().__init__()
        self.concat_xyz_pts = concat_xyz_pts
        self.mlp = Sequential(
            Linear(hidden_size + int(concat_xyz_pts) * 3, hidden_size), LeakyReLU(), Linear(hidden_size, output_dim)
        )
        self.hidden_size = hidden_size
        self.output_dim = output_dim

    def packed(self, device):
        l0, l2 = self.mlp[0], self.mlp[2]
        t = lambda w: w.detach().to(device, F32).t().contiguous()
        v = lambda b: b.detach().to(device, F32).contiguous()
        return t(l0.weight), v(l0.bias), t(l2.weight), v(l2.bias)

    def run(self, vols_cl: List[torch.Tensor], C0: int, vg: VirtualGrid, query: torch.Tensor, emb=None, temperature=1.0):
        """vols_cl: one or two channels-last volumes [N,X,Y,Z,C0]; query [N,nq,3] -> [N,nq,out] (or [N,nq])."""
        dev = query.device
        N, nq = query.shape[:2]
        w1t, b1, w2t, b2 = self.packed(dev)
        out = torch.empty((N, nq) if emb is not None else (N, nq, self.output_dim), device=dev)
        ops.sample_decode(vols_cl[0], vols_cl[1] if len(vols_cl) > 1 else None, C0, query.contiguous().float(),
                          vg.kernel_grid(), N=N, nq=nq, concat_xyz=self.concat_xyz_pts, w1t=w1t, b1=b1, w2t=w2t, b2=b2,
                          Hs=self.hidden_size, out_dim=self.output_dim, out=out, emb=emb, temperature=temperature)
        return out


class PointingAttention(Module):
    def __init__(self, pointing_dim, method="dot_product", pointing_temperature=0.07):
        super().__init__()
        if method != "cosine_sim":
            raise NotImplementedError("only pointing_method='cosine_sim' (what train_vool.py uses) is on the hot path")
        self.method = method
        self.pointing_dim = pointing_dim
        self.cosine_sim_temp = pointing_temperature


class SemAbs3D(Module):
    def __init__(self, voxel_shape: Tuple[int, int, int], scene_bounds, unet_num_channels: int, unet_f_maps: int,
                 unet_num_groups: int, unet_num_levels: int, network_inputs: List[str], use_pts_feat_extractor: bool,
                 pts_feat_extractor_hidden_dim: int, reduce_method: str, output_dim=1, device: str = "cuda",
                 decoder_concat_xyz_pts: bool = False, precise: bool = True, **kwargs):
        super().__init__()
        self.device = device
        self.vg = VirtualGrid(scene_bounds=np.array(scene_bounds), batch_size=kwargs["batch_size"],
                              grid_shape=voxel_shape, device=torch.device(device))
        self.register_buffer("steps", torch.zeros(1))
        self.network_inputs = network_inputs
        self.use_pts_feat_extractor = use_pts_feat_extractor
        self.reduce_method = reduce_method
        if "tsdf" in network_inputs:
            raise NotImplementedError("the TSDF input channel is outside the default hot path (utils.py:95-100)")
        self.pts_feature_dim = ("saliency" in network_inputs) + ("rgb" in network_inputs) * 3 + ("patch_masks" in network_inputs)
        vol_in = self.pts_feature_dim
        if use_pts_feat_extractor:
            h = pts_feat_extractor_hidden_dim
            self.pts_feat_extractor = Sequential(
                Linear(self.pts_feature_dim + 3, h), LeakyReLU(), Linear(h, h), LeakyReLU(), Linear(h, unet_num_channels)
            )
            vol_in = unet_num_channels
            assert self.reduce_method == "max"  # reference assertion (net.py:369); the op it runs is still mean
        self.vol_feature_extractor = ResidualUNet3D(in_channels=vol_in, out_channels=unet_num_channels, f_maps=unet_f_maps,
                                                    num_groups=unet_num_groups, num_levels=unet_num_levels, precise=precise)
        self.visual_sampler = ImplicitVolumetricDecoder(hidden_size=unet_num_channels, output_dim=output_dim,
                                                        concat_xyz_pts=decoder_concat_xyz_pts)
        self.unet_num_channels = unet_num_channels
        self._vol_cl = None
        self._vol_meta = None

    # the reference keeps the UNet output as an attribute in NCDHW (net.py:425-427); converted on demand
    @property
    def visual_volumetric_features(self):
        if self._vol_cl is None:
            return None
        N, (X, Y, Z), C = self._vol_meta
        y = torch.empty(N, C, X, Y, Z, device=self._vol_cl.device)
        ops.ndhwc_to_ncdhw(self._vol_cl, y, N=N, S=X * Y * Z, C=C)
        return y

    def feature_volume(self, input_xyz_pts, input_feature_pts):
        """Everything up to (and including) the UNet; returns the channels-last volume [N, X*Y*Z, C] (a workspace
        buffer that the next call overwrites)."""
        if not input_xyz_pts.is_cuda:
            raise RuntimeError("semabs_b200.SemAbs3D runs on CUDA devices only; there is no CPU path")
        dev = input_xyz_pts.device
        B, P, npts = input_feature_pts.shape[:3]
        N = B * P
        F = self.pts_feature_dim
        unet = self.vol_feature_extractor
        shape = self.vg.grid_shape
        S = shape[0] * shape[1] * shape[2]
        c_in = unet.in_channels
        cpad = _pad16(c_in)
        mlp = None
        if self.use_pts_feat_extractor:
            l0, l2, l4 = self.pts_feat_extractor[0], self.pts_feat_extractor[2], self.pts_feat_extractor[4]
            t = lambda w: w.detach().to(dev, F32).t().contiguous()
            v = lambda b: b.detach().to(dev, F32).contiguous()
            mlp = (t(l0.weight), v(l0.bias), t(l2.weight), v(l2.bias), t(l4.weight), v(l4.bias), l0.out_features)
        vol = unet._buf("l0_in", (N, S, cpad), F32, dev)
        cnt = unet._buf("vox_cnt", (N, S), F32, dev)
        st = unet._buf("l0_pst", (N, 8, 2), F64, dev)
        st.zero_()
        g_in = unet.encoders[0].basic_module.conv1.num_groups
        ops.points_to_voxels(input_xyz_pts.contiguous().float(), input_feature_pts.contiguous().float().view(N, npts, F),
                             self.vg.kernel_grid(), N=N, npts=npts, F=F, xyz_div=P, mlp=mlp, C_out=c_in, vol=vol, cnt=cnt,
                             Cpad=cpad, groups=g_in, stats=st)
        out = unet.forward_channels_last(vol, st, N, shape, dev)
        self._vol_cl, self._vol_meta = out, (N, shape, unet.out_channels)
        return out

    def forward(self, input_xyz_pts, input_feature_pts, tsdf_vol, output_xyz_pts, **kwargs):
        B, P = input_feature_pts.shape[:2]
        nq = output_xyz_pts.shape[2]
        vol = self.feature_volume(input_xyz_pts, input_feature_pts)
        out = self.visual_sampler.run([vol], self.unet_num_channels, self.vg, output_xyz_pts.view(B * P, nq, 3))
        return out.view(B, P, nq, -1).squeeze(dim=-1)


class SemAbsVOOL(Module):
    RELATIONS = ["in", "behind", "in front of", "on the left of", "on the right of", "on", "[pad]"]

    def __init__(self, pointing_method: str, pointing_dim: int, device: str, decoder_concat_xyz_pts: bool, **kwargs):
        super().__init__()
        self.register_buffer("steps", torch.zeros(1))
        self.device = device
        self.completion_net = SemAbs3D(device=device, **kwargs).to(device)
        self.spatial_sampler = ImplicitVolumetricDecoder(hidden_size=2 * kwargs["unet_num_channels"],
                                                         output_dim=pointing_dim, concat_xyz_pts=decoder_concat_xyz_pts)
        self.pointer = PointingAttention(method=pointing_method, pointing_dim=pointing_dim)
        self.relation_embeddings = ParameterDict({k: Parameter(torch.randn(pointing_dim)) for k in self.RELATIONS})

    def get_region_pointing_features(self, spatial_relation_name, **kwargs):
        # [num_descs][batch] names -> [batch, num_descs, pointing_dim] (net.py:505-526); a gather of parameter rows
        num_descs, B = len(spatial_relation_name), len(spatial_relation_name[0])
        rows = [self.relation_embeddings[spatial_relation_name[d][b]] for b in range(B) for d in range(num_descs)]
        return torch.stack(rows, dim=0).view(B, num_descs, -1)

    def forward(self, output_xyz_pts, spatial_relation_name, input_xyz_pts, input_target_saliency_pts,
                input_reference_saliency_pts, tsdf_vol=None, **kwargs):
        B, num_descs = np.array(spatial_relation_name).T.shape
        net = self.completion_net
        C = net.unet_num_channels
        # the UNet output buffer is a workspace: copy the first volume before the second pass overwrites it
        tgt = net.feature_volume(input_xyz_pts, input_target_saliency_pts).clone()
        ref = net.feature_volume(input_xyz_pts, input_reference_saliency_pts)
        nq = output_xyz_pts.shape[-2]
        emb = self.get_region_pointing_features(spatial_relation_name).reshape(B * num_descs, -1).detach().float().contiguous()
        out = self.spatial_sampler.run([tgt, ref], C, net.vg, output_xyz_pts.reshape(B * num_descs, nq, 3), emb=emb,
                                       temperature=self.pointer.cosine_sim_temp)
        return out.view(B, num_descs, nq)
