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


def _ones(n, dev, _cache={}):
    t = _cache.get((n, str(dev)))
    if t is None:
        t = _cache[(n, str(dev))] = torch.ones(n, 1, device=dev)
    return t


def _linear_grads(scratch, P, delta_off, n_out, in_off, n_in, dev):
    """nn.Linear parameter gradients from scratch rows: dW[o][i] = sum_p delta[p][o] * in[p][i], db[o] = sum_p delta[p][o]."""
    dW = torch.zeros(n_out, n_in, device=dev)
    db = torch.zeros(1, n_out, device=dev)
    ops.outer_reduce_f32(scratch[:, delta_off:], scratch[:, in_off:], dW, R=n_out, Cc=n_in, P=P)
    ops.outer_reduce_f32(_ones(P, dev), scratch[:, delta_off:], db, R=1, Cc=n_out, P=P)
    return dW, db.view(n_out)


def _pad4(n):
    return (n + 3) // 4 * 4


class _SampleDecodeFn(torch.autograd.Function):
    """ImplicitVolumetricDecoder.forward (+ cosine pointing head) with its hand-written backward (points_bwd.cu)."""

    @staticmethod
    def forward(ctx, vol0, vol1, query, emb, dec, vg, C0, temperature, w1, b1, w2, b2):
        vols = [vol0] if vol1 is None else [vol0, vol1]
        out = dec._run(vols, C0, vg, query, emb.detach() if emb is not None else None, temperature)
        ctx.save_for_backward(vol0, vol1, query, emb)
        ctx.dec, ctx.vg, ctx.C0, ctx.temperature = dec, vg, C0, temperature
        return out

    @staticmethod
    def backward(ctx, dout):
        vol0, vol1, query, emb = ctx.saved_tensors
        dec, vg, C0 = ctx.dec, ctx.vg, ctx.C0
        dev = query.device
        N, nq = query.shape[:2]
        nvol = 1 if vol1 is None else 2
        Cin = C0 * nvol + (3 if dec.concat_xyz_pts else 0)
        Hs, od = dec.hidden_size, dec.output_dim
        off_h = _pad4(Cin)
        off_do = off_h + _pad4(Hs)
        off_dp = off_do + _pad4(od)
        ld = off_dp + _pad4(Hs)
        P = N * nq
        scratch = torch.empty(P, ld, device=dev)
        l0, l2 = dec.mlp[0], dec.mlp[2]
        w1t, b1, w2t, b2 = dec.packed(dev)
        dvol0 = torch.zeros_like(vol0) if ctx.needs_input_grad[0] else None
        dvol1 = torch.zeros_like(vol1) if (vol1 is not None and ctx.needs_input_grad[1]) else None
        demb = torch.zeros_like(emb) if (emb is not None and ctx.needs_input_grad[3]) else None
        ops.sample_decode_bwd(vol0, vol1, C0, query.contiguous().float(), vg.kernel_grid(), N=N, nq=nq,
                              concat_xyz=dec.concat_xyz_pts, w1t=w1t, w1=l0.weight.detach().float().contiguous(), b1=b1,
                              w2t=w2t, w2=l2.weight.detach().float().contiguous(), b2=b2, Hs=Hs, out_dim=od,
                              dout=dout.contiguous().float(), dvol0=dvol0, dvol1=dvol1, scratch=scratch, off_h=off_h,
                              off_do=off_do, off_dp=off_dp, emb=emb.detach() if emb is not None else None,
                              temperature=ctx.temperature, demb=demb)
        dW2, db2 = _linear_grads(scratch, P, off_do, od, off_h, Hs, dev)
        dW1, db1 = _linear_grads(scratch, P, off_dp, Hs, 0, Cin, dev)
        return dvol0, dvol1, None, demb, None, None, None, None, dW1, db1, dW2, db2


class _FeatureVolumeFn(torch.autograd.Function):
    """points -> [point MLP + scatter-mean] -> ResidualUNet3D, channels-last output [N, X*Y*Z, C]; backward = UNet
    backward (unet3d_bwd.py) followed by the scatter-mean / point-MLP backward (points_bwd.cu)."""

    @staticmethod
    def forward(ctx, xyz, feat, net, *params):
        unet = net.vol_feature_extractor
        tape = unet._bwd().new_tape()
        out = net._feature_volume(xyz, feat, tape=tape)
        ctx.net, ctx.tape, ctx.params = net, tape, params
        ctx.save_for_backward(xyz, feat)
        return out

    @staticmethod
    def backward(ctx, d_out):
        net, tape = ctx.net, ctx.tape
        xyz, feat = ctx.saved_tensors
        unet = net.vol_feature_extractor
        bw = unet._bwd()
        dev = d_out.device
        from .train import GradientBuckets

        buckets = GradientBuckets.active()  # data-parallel training: per-level all-reduce overlapped with this backward
        dx_cl, grads = bw.backward(tape, d_out.contiguous().float(), need_dx=net.use_pts_feat_extractor, buckets=buckets)
        n_unet = len(grads)
        if net.use_pts_feat_extractor:
            ex = tape.extra
            N, npts, F, P, cpad = ex["N"], ex["npts"], ex["F"], ex["P"], ex["cpad"]
            l0, l2, l4 = net.pts_feat_extractor[0], net.pts_feat_extractor[2], net.pts_feat_extractor[4]
            Hd, C = l0.out_features, l4.out_features
            w1t, b1, w2t, b2, _, _, _ = net._mlp_pack(dev)
            off_d3 = 8 + 2 * Hd
            off_d2 = off_d3 + _pad4(C)
            off_d1 = off_d2 + Hd
            ld = off_d1 + Hd
            npt = N * npts
            scratch = torch.empty(npt, ld, device=dev)
            ops.points_to_voxels_bwd(xyz.contiguous().float(), feat.contiguous().float().view(N, npts, F), net.vg.kernel_grid(),
                                     N=N, npts=npts, F=F, xyz_div=P, hidden=Hd, C=C, w1t=w1t, b1=b1, w2t=w2t,
                                     w2=l2.weight.detach().float().contiguous(), b2=b2,
                                     w3=l4.weight.detach().float().contiguous(), dvol=dx_cl, cnt=ex["cnt"], Cpad=cpad,
                                     scratch=scratch, off_d3=off_d3, off_d2=off_d2, off_d1=off_d1)
            grads[l4.weight], grads[l4.bias] = _linear_grads(scratch, npt, off_d3, C, 8 + Hd, Hd, dev)
            grads[l2.weight], grads[l2.bias] = _linear_grads(scratch, npt, off_d2, Hd, 8, Hd, dev)
            grads[l0.weight], grads[l0.bias] = _linear_grads(scratch, npt, off_d1, Hd, 0, 3 + F, dev)
        bw.release(tape)
        if buckets is not None:
            buckets.submit(grads, list(grads)[n_unet:])  # the point MLP: one last small bucket
            buckets.join(grads)
        return (None, None, None) + tuple(grads.get(p) for p in ctx.params)


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
        if torch.is_grad_enabled() and (any(v.requires_grad for v in vols_cl) or any(p.requires_grad for p in self.parameters())
                                        or (emb is not None and emb.requires_grad)):
            l0, l2 = self.mlp[0], self.mlp[2]
            return _SampleDecodeFn.apply(vols_cl[0], vols_cl[1] if len(vols_cl) > 1 else None, query, emb, self, vg, C0,
                                         float(temperature), l0.weight, l0.bias, l2.weight, l2.bias)
        return self._run(vols_cl, C0, vg, query, emb, temperature)

    def _run(self, vols_cl, C0, vg, query, emb=None, temperature=1.0):
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
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            params = list(self.pts_feat_extractor.parameters()) if self.use_pts_feat_extractor else []
            params += list(self.vol_feature_extractor.parameters())
            out = _FeatureVolumeFn.apply(input_xyz_pts, input_feature_pts, self, *params)
            B, P = input_feature_pts.shape[:2]
            self._vol_cl, self._vol_meta = out.detach(), (B * P, self.vg.grid_shape, self.vol_feature_extractor.out_channels)
            return out
        return self._feature_volume(input_xyz_pts, input_feature_pts)

    def _mlp_pack(self, dev):
        l0, l2, l4 = self.pts_feat_extractor[0], self.pts_feat_extractor[2], self.pts_feat_extractor[4]
        t = lambda w: w.detach().to(dev, F32).t().contiguous()
        v = lambda b: b.detach().to(dev, F32).contiguous()
        return (t(l0.weight), v(l0.bias), t(l2.weight), v(l2.bias), t(l4.weight), v(l4.bias), l0.out_features)

    def _feature_volume(self, input_xyz_pts, input_feature_pts, tape=None):
        dev = input_xyz_pts.device
        B, P, npts = input_feature_pts.shape[:3]
        N = B * P
        F = self.pts_feature_dim
        unet = self.vol_feature_extractor
        shape = self.vg.grid_shape
        S = shape[0] * shape[1] * shape[2]
        c_in = unet.in_channels
        cpad = _pad16(c_in)
        mlp = self._mlp_pack(dev) if self.use_pts_feat_extractor else None
        vol = unet._alloc(tape, "l0_in", "l0_in", (N, S, cpad), F32, dev)
        cnt = unet._alloc(tape, "vox_cnt", "vox_cnt", (N, S), F32, dev)
        st = unet._alloc(tape, "l0_pst", "l0_pst", (N, 8, 2), F64, dev)
        st.zero_()
        g_in = unet.encoders[0].basic_module.conv1.num_groups
        ops.points_to_voxels(input_xyz_pts.contiguous().float(), input_feature_pts.contiguous().float().view(N, npts, F),
                             self.vg.kernel_grid(), N=N, npts=npts, F=F, xyz_div=P, mlp=mlp, C_out=c_in, vol=vol, cnt=cnt,
                             Cpad=cpad, groups=g_in, stats=st)
        out = unet.forward_channels_last(vol, st, N, shape, dev, tape=tape)
        if tape is not None:
            tape.extra.update(cnt=cnt, N=N, npts=npts, F=F, P=P, cpad=cpad)
        else:
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
        tgt = net.feature_volume(input_xyz_pts, input_target_saliency_pts)
        if not tgt.requires_grad:
            tgt = tgt.clone()  # inference: the UNet output is a workspace the second pass overwrites
        ref = net.feature_volume(input_xyz_pts, input_reference_saliency_pts)
        nq = output_xyz_pts.shape[-2]
        emb = self.get_region_pointing_features(spatial_relation_name).reshape(B * num_descs, -1).float().contiguous()
        if not torch.is_grad_enabled():
            emb = emb.detach()
        out = self.spatial_sampler.run([tgt, ref], C, net.vg, output_xyz_pts.reshape(B * num_descs, nq, 3), emb=emb,
                                       temperature=self.pointer.cosine_sim_temp)
        return out.view(B, num_descs, nq)
