"""ResidualUNet3D on the B200 kernels — host-side mirror of the reference module of the same name
(reference: unet3d.py:658-689 on top of Abstract3DUNet :481-621, ExtResNetBlock :190-259, Encoder :262-317,
Decoder/Upsampling :320-444).

The nn.Module tree exists only to own parameters under the reference's state-dict key names
(`encoders.0.basic_module.conv1.groupnorm.weight`, `decoders.0.upsampling.upsample.weight`, `final_conv.bias`, ...)
and — being built in the reference's construction order from stock torch containers — to draw identical
initial values under the same torch seed.  forward() never calls those containers: it issues the C-ABI kernels
(GroupNorm-apply -> implicit-GEMM conv with fused ReLU / residual / next-GroupNorm statistics; max-pool with fused
statistics; transposed conv as 8 parity classes with fused skip-sum), on channels-last buffers.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import os

import torch
import torch.nn as nn

from . import ops

F16, F32, F64 = torch.float16, torch.float32, torch.float64


_WEIGHTS_EPOCH = [0]


def bump_weights_epoch() -> None:
    """Called by optimisers that update parameters through raw device pointers (train.Lamb): such updates do not move
    torch's tensor version counters, so the cached fp16 weight packs are additionally keyed on this epoch."""
    _WEIGHTS_EPOCH[0] += 1


def _params_key(module, device, precise):
    ps = list(module.parameters())
    return (str(device), precise, _WEIGHTS_EPOCH[0], tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))


def number_of_features_per_level(init_channel_number, num_levels):
    return [init_channel_number * 2**k for k in range(num_levels)]


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


class SingleConv(nn.Sequential):
    """'gc[r]' unit: GroupNorm(in) -> Conv3d(3x3x3, pad 1, no bias) [-> ReLU] (create_conv, unet3d.py:20-95)."""

    def __init__(self, in_channels, out_channels, order="gcr", num_groups=8):
        super().__init__()
        assert order in ("gcr", "gc"), "only the reference's ResidualUNet3D layer order 'gcr' is implemented"
        if in_channels < num_groups:
            num_groups = 1  # unet3d.py:72-73
        self.add_module("groupnorm", nn.GroupNorm(num_groups=num_groups, num_channels=in_channels))
        self.add_module("conv", nn.Conv3d(in_channels, out_channels, 3, padding=1, bias=False))
        self.relu = "r" in order
        self.num_groups = num_groups


class ExtResNetBlock(nn.Module):
    def __init__(self, in_channels, out_channels, order="gcr", num_groups=8, **kwargs):
        super().__init__()
        self.conv1 = SingleConv(in_channels, out_channels, order, num_groups)
        self.conv2 = SingleConv(out_channels, out_channels, order, num_groups)
        self.conv3 = SingleConv(out_channels, out_channels, order.replace("r", ""), num_groups)


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, apply_pooling=True, conv_layer_order="gcr", num_groups=8):
        super().__init__()
        self.pooling = nn.MaxPool3d(kernel_size=(2, 2, 2)) if apply_pooling else None
        self.basic_module = ExtResNetBlock(in_channels, out_channels, order=conv_layer_order, num_groups=num_groups)


class Upsampling(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.upsample = nn.ConvTranspose3d(in_channels, out_channels, kernel_size=3, stride=(2, 2, 2), padding=1)


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, conv_layer_order="gcr", num_groups=8):
        super().__init__()
        self.upsampling = Upsampling(in_channels, out_channels)
        self.basic_module = ExtResNetBlock(out_channels, out_channels, order=conv_layer_order, num_groups=num_groups)


def _split_pack(w2d: torch.Tensor, splits: int) -> torch.Tensor:
    """[Co, K] fp32 -> [Co, splits*K] fp16 (hi | lo)."""
    hi = w2d.half()
    if splits == 1:
        return hi.contiguous()
    lo = (w2d - hi.float()).half()
    return torch.cat([hi, lo], dim=1).contiguous()


class ResidualUNet3D(nn.Module):
    """Same constructor arguments as the reference. `precise=True` (default) runs every convolution as
    x_hi·w_hi + x_lo·w_hi + x_hi·w_lo on fp16 tensor-core MMAs with fp32 accumulation (≈ fp32 accuracy: the voxel
    logits have to match the fp32 reference to 1e-3); `precise=False` is the single-pass fp16 mode."""

    def __init__(self, in_channels, out_channels, f_maps=64, num_groups=8, num_levels=5, final_sigmoid=False,
                 layer_order="gcr", is_segmentation=False, precise=True, **kwargs):
        super().__init__()
        if isinstance(f_maps, int):
            f_maps = number_of_features_per_level(f_maps, num_levels=num_levels)
        self.f_maps = list(f_maps)
        self.in_channels, self.out_channels, self.num_groups = in_channels, out_channels, num_groups
        self.precise = precise
        self.use_halo = True  # halo-resident conv kernel at the 128-wide level (conv3d_halo.cu)
        # inference, 128-wide level, 32 channels, precise mode: conv2 / conv3 of a block consume the RAW output of their producer
        # (written chunk-planar hi | lo by its epilogue) with the GroupNorm folded into per-sample weights + a border-class bias
        # (semabs_conv3d_halo_fused) instead of a GroupNorm-apply pass over an fp32 copy; the fold is one kernel launch per
        # convolution for all samples (semabs_fold_groupnorm_halo)
        self.fold_groupnorm = True
        self.folded_blocks = 0  # residual blocks that took the folded path (tests)
        self.fuse_transposed = True  # transposed convolution into a 32-channel level: eight parity classes in one launch
        # inference forward replayed as a CUDA graph (see _forward_graphed); SEMABS_UNET_GRAPH=0 = eager launches (A/B runs)
        self.use_cuda_graph = os.environ.get("SEMABS_UNET_GRAPH", "1") != "0"
        self._graphs: Dict[Tuple, tuple] = {}
        self._graph_seen: set = set()
        self._pack_gen = 0
        encoders = []
        for i, out_f in enumerate(f_maps):
            encoders.append(Encoder(in_channels if i == 0 else f_maps[i - 1], out_f, apply_pooling=i > 0,
                                    conv_layer_order=layer_order, num_groups=num_groups))
        self.encoders = nn.ModuleList(encoders)
        rev = list(reversed(f_maps))
        self.decoders = nn.ModuleList(
            [Decoder(rev[i], rev[i + 1], conv_layer_order=layer_order, num_groups=num_groups) for i in range(len(rev) - 1)]
        )
        self.final_conv = nn.Conv3d(f_maps[0], out_channels, 1)
        assert not is_segmentation, "final activation is not used on the hot path (reference default is_segmentation=False)"
        self._pack: Dict[str, torch.Tensor] = {}
        self._pack_key = None
        self._ws: Dict[Tuple, torch.Tensor] = {}
        self._bwd_engine = None
        self.kernel_launches = 0

    # ------------------------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype, device):
        key = (name, tuple(int(s) for s in shape), dtype, str(device))
        t = self._ws.get(key)
        if t is None:
            for k in [k for k in self._ws if k[0] == name]:
                del self._ws[k]
            t = torch.empty(key[1], dtype=dtype, device=device)
            self._ws[key] = t
            self._graphs.clear()  # captured launches hold workspace addresses: a replaced buffer invalidates every graph
            self._graph_seen.clear()
        return t

    def _bwd(self):
        if getattr(self, "_bwd_engine", None) is None:
            from .unet3d_bwd import UNetBackward

            self._bwd_engine = UNetBackward(self)
        return self._bwd_engine

    def _alloc(self, tape, ws_name, tape_name, shape, dtype, dev):
        """inference: a reusable workspace buffer; training (tape given): a buffer that survives until backward"""
        if tape is None:
            return self._buf(ws_name, shape, dtype, dev)
        return tape.alloc(tape_name, shape, dtype, dev)

    def _packed(self, device):
        """fp16 (hi | lo) MMA-operand copies of the conv weights, rebuilt when a parameter changes."""
        key = _params_key(self, device, self.precise)
        if key == self._pack_key:
            return self._pack
        s = 2 if self.precise else 1
        pk: Dict[str, torch.Tensor] = {}

        def conv_w(w):  # [Co, Ci, 3,3,3] -> [Co, s*27*Ci_pad]
            co, ci = w.shape[:2]
            wp = torch.zeros(co, 27, _pad16(ci), device=device)
            wp[:, :, :ci] = w.detach().to(device, F32).permute(0, 2, 3, 4, 1).reshape(co, 27, ci)
            return _split_pack(wp.reshape(co, -1), s)

        def gn(m, cpad):
            g = torch.zeros(cpad, device=device)
            b = torch.zeros(cpad, device=device)
            g[: m.num_channels] = m.weight.detach().to(device, F32)
            b[: m.num_channels] = m.bias.detach().to(device, F32)
            return g, b

        def halo_w(w):  # [Co, Ci, 3,3,3] -> per-tap core-matrix images for the halo-resident kernel (W == 128 level)
            co, ci = w.shape[:2]
            wp = torch.zeros(co, _pad16(ci), 3, 3, 3, device=device)
            wp[:, :ci] = w.detach().to(device, F32)
            return ops.pack_halo_weights(wp, s)

        def block(prefix, blk: ExtResNetBlock):
            for j, sc in enumerate((blk.conv1, blk.conv2, blk.conv3), 1):
                pk[f"{prefix}.w{j}"] = conv_w(sc.conv.weight)
                co, ci = sc.conv.weight.shape[:2]
                if co in (16, 32) and _pad16(ci) in (16, 32):
                    pk[f"{prefix}.wh{j}"] = halo_w(sc.conv.weight)
                    if co == 32 and ci == 32:
                        pk[f"{prefix}.wraw{j}"] = sc.conv.weight.detach().to(device, F32).contiguous()
                pk[f"{prefix}.g{j}"], pk[f"{prefix}.b{j}"] = gn(sc.groupnorm, _pad16(sc.groupnorm.num_channels))

        for i, enc in enumerate(self.encoders):
            block(f"enc{i}", enc.basic_module)
        for i, dec in enumerate(self.decoders):
            block(f"dec{i}", dec.basic_module)
            w = dec.upsampling.upsample.weight  # [Ci, Co, 3,3,3]
            ci, co = w.shape[:2]
            pk[f"dec{i}.up_w"] = _split_pack(w.detach().to(device, F32).permute(1, 2, 3, 4, 0).reshape(co, 27 * ci), s)
            pk[f"dec{i}.up_b"] = dec.upsampling.upsample.bias.detach().to(device, F32).contiguous()
        fw = self.final_conv.weight
        pk["final.w"] = _split_pack(fw.detach().to(device, F32).reshape(fw.shape[0], fw.shape[1]), s)
        pk["final.b"] = self.final_conv.bias.detach().to(device, F32).contiguous()
        pk["final.w32"] = fw.detach().to(device, F32).reshape(fw.shape[0], fw.shape[1]).contiguous()
        self._pack, self._pack_key = pk, key
        self._pack_gen += 1
        return pk

    # ------------------------------------------------------------------------------------------------
    def _res_block(self, pk, prefix, blk: ExtResNetBlock, x_raw, x_stats, *, N, dims, c_in_pad, c_in_real, lvl, dev,
                   want32: bool, want16: bool, tape=None, x_planar=None):
        """ExtResNetBlock.forward (unet3d.py:243-259): o1 = relu(conv(gn(x))); o2 = relu(conv(gn(o1)));
        out = relu(conv(gn(o2)) + o1). Returns (out32 | None, out16 | None)."""
        D, H, W = dims
        S = D * H * W
        s = 2 if self.precise else 1
        c_out = blk.conv1.conv.out_channels
        if self._folds(pk, prefix, blk, tape, dims, c_in_pad):
            return self._res_block_folded(pk, prefix, blk, x_raw, x_stats, N=N, dims=dims, c_in_pad=c_in_pad, c_in_real=c_in_real,
                                          lvl=lvl, dev=dev, want32=want32, want16=want16, x_planar=x_planar)
        assert x_planar is None, "a planar raw input is only produced for blocks that fold their first GroupNorm"
        xn = self._buf(f"l{lvl}_xn", (N, S, s * max(c_in_pad, c_out)), F16, dev)
        o1 = self._alloc(tape, f"l{lvl}_o1", f"{prefix}.o1", (N, S, c_out), F32, dev)
        o2 = self._alloc(tape, f"l{lvl}_o2", f"{prefix}.o2", (N, S, c_out), F32, dev)
        g2, g3 = blk.conv2.num_groups, blk.conv3.num_groups
        st = self._alloc(tape, f"l{lvl}_st", f"{prefix}.st", (2, N, 8, 2), F64, dev)
        st.zero_()
        want32 = want32 or tape is not None  # the backward needs every block output (ReLU mask, next block's input)
        halo = W == 128 and c_in_pad in (16, 32) and c_out in (16, 32) and self.use_halo
        common = dict(N=N, D=D, H=H, W=W, a_splits=s, w_splits=s, precise=self.precise)

        def gcr(j, src, src_stats, c_src_pad, c_src_real, gn_groups, **epi):
            """one 'gc[r]' unit: GroupNorm-apply then the implicit-GEMM conv (halo-resident kernel at full resolution)"""
            ops.groupnorm_apply(src, src_stats, pk[f"{prefix}.g{j}"], pk[f"{prefix}.b{j}"], xn, N=N, S=S, C=c_src_pad,
                                C_real=c_src_real, groups=gn_groups, splits=s, planar=halo)
            if halo:
                ops.conv3d_halo(xn, pk[f"{prefix}.wh{j}"], C_in=c_src_pad, C_out=c_out, **common, **epi)
            else:
                ops.conv3d(xn, pk[f"{prefix}.w{j}"], kind=ops.CONV_3X3X3, C_in=c_src_pad, C_out=c_out, **common, **epi)

        gcr(1, x_raw, x_stats, c_in_pad, c_in_real, blk.conv1.num_groups, relu=True, out32=o1, stats=st[0], groups=g2)
        gcr(2, o1, st[0], c_out, c_out, g2, relu=True, out32=o2, stats=st[1], groups=g3)
        out32 = self._alloc(tape, f"l{lvl}_out32", f"{prefix}.out", (N, S, c_out), F32, dev) if want32 else None
        out16 = self._buf(f"l{lvl}_out16", (N, S, s * c_out), F16, dev) if want16 else None
        gcr(3, o2, st[1], c_out, c_out, g3, residual=o1, relu=True, out32=out32, out16=out16, o16_splits=s)
        self.kernel_launches += 6
        if tape is not None:
            tape.blocks[prefix] = dict(x=x_raw, x_stats=x_stats, o1=o1, o2=o2, st=st, out=out32, dims=dims,
                                       c_in_pad=c_in_pad, c_in_real=c_in_real, c_out=c_out)
        return out32, out16


    def _folds(self, pk, prefix, blk, tape, dims, c_in_pad) -> bool:
        """This residual block runs with its GroupNorms folded into the halo convolutions (inference, 128-wide level, 32 channels)."""
        D, H, W = dims
        return (tape is None and self.fold_groupnorm and self.use_halo and self.precise and W == 128 and H % 2 == 0
                and blk.conv1.conv.out_channels == 32 and c_in_pad in (16, 32) and blk.conv2.num_groups * 2 <= 32
                and f"{prefix}.wraw2" in pk)

    def _folds_input(self, pk, prefix, blk, tape, dims, c_in_pad, c_in_real) -> bool:
        """... and also its FIRST GroupNorm: the producer then writes the raw block input chunk-planar hi | lo itself."""
        return (self._folds(pk, prefix, blk, tape, dims, c_in_pad) and c_in_pad == 32 and c_in_real == 32 and f"{prefix}.wraw1" in pk
                and (32 // blk.conv1.num_groups) % 4 == 0)

    def _fold_gn(self, key, w, gamma, beta, stats, S, groups):
        """GroupNorm(x) = a x + b per (sample, channel) with a = gamma rstd, b = beta - mean a, so conv(GroupNorm(x)) =
        conv_{W a}(x) + sum over the taps that fall inside the grid of (W b): -> (weight images [N, ...] of W a,
        bias table [N, 27 border classes, C_out]), one kernel launch for all samples (semabs_fold_groupnorm_halo)."""
        N = stats.shape[0]
        imgs = self._buf(f"{key}_img", (N, 55296), F16, w.device)
        bias = self._buf(f"{key}_bias", (N, 27, 32), F32, w.device)
        ops.fold_groupnorm_halo(w, gamma, beta, stats, imgs, bias, N=N, S=S, groups=groups)
        self.kernel_launches += 1
        return imgs, bias

    def _res_block_folded(self, pk, prefix, blk, x_raw, x_stats, *, N, dims, c_in_pad, c_in_real, lvl, dev, want32, want16,
                          x_planar=None):
        """ExtResNetBlock.forward at the 128-wide level with 32 channels, inference: conv2 / conv3 read the raw planar output of
        conv1 / conv2 with folded GroupNorm; o1 / o2 exist only as chunk-planar hi | lo fp16, the residual of conv3 is read from
        o1's planar copy.  conv1 reads either a GroupNorm-applied copy of x_raw (fp32 channels-last) or, when its producer wrote
        the raw input chunk-planar itself (x_planar), that tensor with the first GroupNorm folded as well."""
        D, H, W = dims
        S = D * H * W
        c = 32
        g2, g3 = blk.conv2.num_groups, blk.conv3.num_groups
        o1p = self._buf(f"l{lvl}_o1p", (N, 2 * c * S), F16, dev)
        o2p = self._buf(f"l{lvl}_o2p", (N, 2 * c * S), F16, dev)
        st = self._buf(f"l{lvl}_st", (2, N, 8, 2), F64, dev)
        st.zero_()
        kw = dict(D=D, H=H, a_splits=2, w_splits=2, precise=True)
        if x_planar is not None:
            w1, bias1 = self._fold_gn(f"l{lvl}_f1", pk[f"{prefix}.wraw1"], pk[f"{prefix}.g1"], pk[f"{prefix}.b1"], x_stats, S,
                                      blk.conv1.num_groups)
            xp = x_planar.view(N, -1)
            for n in range(N):
                ops.conv3d_halo_fused(xp[n], w1[n], C_in=c, bias_cls=bias1[n], relu=True, out_planar=o1p[n], stats=st[0, n], groups=g2, **kw)
        else:
            xn = self._buf(f"l{lvl}_xn", (N, S, 2 * max(c_in_pad, c)), F16, dev)
            ops.groupnorm_apply(x_raw, x_stats, pk[f"{prefix}.g1"], pk[f"{prefix}.b1"], xn, N=N, S=S, C=c_in_pad, C_real=c_in_real,
                                groups=blk.conv1.num_groups, splits=2, planar=True)
            xn_n = xn.view(N, -1)
            for n in range(N):
                ops.conv3d_halo_fused(xn_n[n], pk[f"{prefix}.wh1"], C_in=c_in_pad, relu=True, out_planar=o1p[n], stats=st[0, n], groups=g2, **kw)
        w2, bias2 = self._fold_gn(f"l{lvl}_f2", pk[f"{prefix}.wraw2"], pk[f"{prefix}.g2"], pk[f"{prefix}.b2"], st[0], S, g2)
        for n in range(N):
            ops.conv3d_halo_fused(o1p[n], w2[n], C_in=c, bias_cls=bias2[n], relu=True, out_planar=o2p[n], stats=st[1, n], groups=g3, **kw)
        w3, bias3 = self._fold_gn(f"l{lvl}_f3", pk[f"{prefix}.wraw3"], pk[f"{prefix}.g3"], pk[f"{prefix}.b3"], st[1], S, g3)
        out32 = self._buf(f"l{lvl}_out32", (N, S, c), F32, dev) if want32 else None
        out16 = self._buf(f"l{lvl}_out16", (N, S, 2 * c), F16, dev) if want16 else None
        for n in range(N):
            ops.conv3d_halo_fused(o2p[n], w3[n], C_in=c, bias_cls=bias3[n], res_planar=o1p[n], relu=True,
                                  out32=out32[n] if want32 else None, out16=out16[n] if want16 else None, o16_splits=2, **kw)
        self.kernel_launches += 1 + 3 * N
        self.folded_blocks += 1
        return out32, out16

    def forward_channels_last(self, x_raw, x_stats, N, dims, dev, tape=None, ncdhw_out=None, x_planar=None):
        """Core of Abstract3DUNet.forward (unet3d.py:596-621) on channels-last buffers. x_raw [N,S,Cpad] fp32 with
        its GroupNorm statistics. Returns the final conv output, channels-last fp32 [N,S,out_channels]."""
        pk = self._packed(dev)
        s = 2 if self.precise else 1
        L = len(self.encoders)
        feats: List[Tuple[torch.Tensor, Tuple[int, int, int]]] = []
        cur_raw, cur_stats = x_raw, x_stats
        c_pad, c_real = _pad16(self.in_channels), self.in_channels
        cur16 = None
        for i, enc in enumerate(self.encoders):
            if i > 0:
                D, H, W = dims
                g_next = enc.basic_module.conv1.num_groups
                pooled = self._alloc(tape, f"l{i}_in", f"enc{i}.in", (N, (D // 2) * (H // 2) * (W // 2), c_pad), F32, dev)
                pst = self._alloc(tape, f"l{i}_pst", f"enc{i}.pst", (N, 8, 2), F64, dev)
                pst.zero_()
                ops.maxpool3d_2(cur_raw, pooled, N=N, D=D, H=H, W=W, C=c_pad, groups=g_next, stats=pst)
                self.kernel_launches += 1
                dims = (D // 2, H // 2, W // 2)
                cur_raw, cur_stats = pooled, pst
            last = i == L - 1
            out32, out16 = self._res_block(pk, f"enc{i}", enc.basic_module, cur_raw, cur_stats, N=N, dims=dims,
                                           c_in_pad=c_pad, c_in_real=c_real, lvl=i, dev=dev, want32=not last,
                                           want16=last, tape=tape, x_planar=x_planar if i == 0 else None)
            c_pad = c_real = self.f_maps[i]
            if not last:
                feats.insert(0, (out32, dims))
                cur_raw = out32
            else:
                cur16 = out16
        for j, (dec, (skip, sdims)) in enumerate(zip(self.decoders, feats)):
            lvl = L - 2 - j
            D, H, W = dims  # input grid of the transposed conv
            c_in, c_out = self.f_maps[lvl + 1], self.f_maps[lvl]
            S_out = sdims[0] * sdims[1] * sdims[2]
            g1 = dec.basic_module.conv1.num_groups
            assert sdims == (2 * D, 2 * H, 2 * W), "ConvTranspose3d(output_size) path expects exact 2x up-sampling"
            fused_t = self.fuse_transposed and c_out in (32, 64) and c_in % 64 == 0 and D * H * W * N >= 32 and (c_out // g1) % 4 == 0
            # the block that consumes the up-sampled tensor folds its first GroupNorm too when the transposed convolution can hand
            # it the raw tensor chunk-planar (hi | lo): then no fp32 copy of it exists at all
            up_planar = None
            if fused_t and (2 * W) % 32 == 0 and self._folds_input(pk, f"dec{j}", dec.basic_module, tape, sdims, c_out, c_out):
                up_planar = self._buf(f"l{lvl}_upp", (N, 2 * c_out * S_out), F16, dev)
                up = None
            else:
                up = self._alloc(tape, f"l{lvl}_up", f"dec{j}.up", (N, S_out, c_out), F32, dev)
            ust = self._alloc(tape, f"l{lvl}_ust", f"dec{j}.ust", (N, 8, 2), F64, dev)
            ust.zero_()
            # Upsampling.forward + summation joining (unet3d.py:385-396, 438-440)
            if fused_t:
                # into the 32- / 64-channel levels (the big grids): all eight output-parity classes in one pass (conv3d_convt.cu)
                ops.conv_transpose3d_s2(cur16, pk[f"dec{j}.up_w"], N=N, D=D, H=H, W=W, C_in=c_in, C_out=c_out, a_splits=s, w_splits=s,
                                        precise=self.precise, bias=pk[f"dec{j}.up_b"], residual=skip, out32=up, out_planar=up_planar,
                                        stats=ust, groups=g1)
                self.kernel_launches += 1
            else:
                for parity in range(8):
                    ops.conv3d(cur16, pk[f"dec{j}.up_w"], kind=ops.CONV_TRANSPOSE_PARITY, parity=parity, N=N, D=D, H=H, W=W,
                               C_in=c_in, C_out=c_out, a_splits=s, w_splits=s, precise=self.precise, bias=pk[f"dec{j}.up_b"],
                               residual=skip, out32=up, stats=ust, groups=g1)
                self.kernel_launches += 8
            dims = sdims
            _, cur16 = self._res_block(pk, f"dec{j}", dec.basic_module, up, ust, N=N, dims=dims, c_in_pad=c_out,
                                       c_in_real=c_out, lvl=lvl, dev=dev, want32=False, want16=True, tape=tape, x_planar=up_planar)
        D, H, W = dims
        if ncdhw_out is not None and tape is None and self.f_maps[0] in (16, 32, 64) and self.out_channels <= 256:
            # inference: final_conv and the conversion back to NCDHW in one pass (no channels-last fp32 copy of the output)
            ops.final_conv1x1_ncdhw(cur16, pk["final.w32"], pk["final.b"], ncdhw_out, N=N, S=D * H * W, C_in=self.f_maps[0],
                                    C_out=self.out_channels, splits=s)
            self.kernel_launches += 1
            return None
        out = self._alloc(tape, "final_cl", "final.out", (N, D * H * W, self.out_channels), F32, dev)
        if tape is not None:
            tape.meta = dict(N=N, dims0=dims, dev=dev)
        ops.conv3d(cur16, pk["final.w"], kind=ops.CONV_1X1X1, N=N, D=D, H=H, W=W, C_in=self.f_maps[0],
                   C_out=self.out_channels, a_splits=s, w_splits=s, precise=self.precise, bias=pk["final.b"], out32=out)
        self.kernel_launches += 1
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, C, D, H, W] fp32 on a CUDA device -> [N, out_channels, D, H, W] fp32 (NCDHW, like the reference)."""
        if not x.is_cuda:
            raise RuntimeError("semabs_b200.ResidualUNet3D runs on CUDA devices only; there is no CPU path")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .unet3d_bwd import _UNetFn  # training: the same kernels + a tape, backward in unet3d_bwd.py

            return _UNetFn.apply(x, self, *self.parameters())
        assert x.shape[1] == self.in_channels
        x = x.contiguous().float()
        from ._lib import CALL_PROFILE

        if self.use_cuda_graph and not CALL_PROFILE.on and not torch.cuda.is_current_stream_capturing():
            return self._forward_graphed(x)
        return self._forward_eager(x)

    def _forward_graphed(self, x: torch.Tensor) -> torch.Tensor:
        """Inference forward as a CUDA-graph replay.  One forward is ~350 launches of 10-2000 us kernels: on a busy host the Python /
        ctypes side (20-40 us per launch) is slower than the GPU (measured: 26 ms of kernels, 40-57 ms per forward in some runs).
        Graphs are keyed on everything a captured launch bakes in: input pointer and shape, the weight pack (rebuilt when a
        parameter changes), the path switches.  Workspaces are the module's cached buffers, so their addresses are stable; the
        output lives in the graph's pool and a copy is returned."""
        dev = x.device
        self._packed(dev)  # (re)build the weight pack outside the capture; its identity is part of the key
        key = (x.data_ptr(), tuple(x.shape), str(dev), self._pack_gen, self.precise, self.use_halo, self.fold_groupnorm, self.fuse_transposed)
        entry = self._graphs.get(key)
        if entry is None:
            if key not in self._graph_seen:
                # first sight of this (input buffer, shape): run eagerly — a caller that passes a fresh tensor every time must not
                # pay a capture per call; the second call with the same buffer captures
                y = self._forward_eager(x)  # (may allocate workspaces, which clears the seen set: add the key afterwards)
                if len(self._graph_seen) >= 16:
                    self._graph_seen.clear()
                self._graph_seen.add(key)
                return y
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            torch.cuda.synchronize(dev)
            self._forward_eager(x)  # allocates every workspace of this shape (a new buffer drops all graphs, see _buf)
            torch.cuda.synchronize(dev)
            l0 = self.kernel_launches
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                y_static = self._forward_eager(x)
            entry = (graph, y_static, x, self.kernel_launches - l0)  # (x kept alive: its address is baked into the graph)
            self._graphs[key] = entry
        entry[0].replay()
        self.kernel_launches += entry[3]
        return entry[1].clone()

    def _forward_eager(self, x: torch.Tensor) -> torch.Tensor:
        N, C, D, H, W = x.shape
        dev = x.device
        S = D * H * W
        cpad = _pad16(C)
        g_in = self.encoders[0].basic_module.conv1.num_groups
        st = self._buf("l0_pst", (N, 8, 2), F64, dev)
        st.zero_()
        raw = xp = None
        pk = self._packed(dev)
        if self._folds_input(pk, "enc0", self.encoders[0].basic_module, None, (D, H, W), cpad, C):
            # the first block folds its first GroupNorm: the module input goes straight into the halo kernel's operand layout
            xp = self._buf("l0_inp", (N, 2 * cpad * S), F16, dev)
            ops.ncdhw_to_planar(x, xp, N=N, S=S, C=C, Cpad=cpad, groups=g_in, stats=st)
        else:
            raw = self._buf("l0_in", (N, S, cpad), F32, dev)
            ops.ncdhw_to_ndhwc(x, raw, N=N, S=S, C=C, Cpad=cpad, groups=g_in, stats=st)
        y = torch.empty(N, self.out_channels, D, H, W, device=dev)
        out_cl = self.forward_channels_last(raw, st, N, (D, H, W), dev, ncdhw_out=y, x_planar=xp)
        if out_cl is not None:
            ops.ndhwc_to_ncdhw(out_cl, y, N=N, S=S, C=self.out_channels)
            self.kernel_launches += 1
        self.kernel_launches += 1
        return y
