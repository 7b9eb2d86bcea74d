"""Backward pass of ResidualUNet3D on the B200 kernels (what `loss.backward()` does through the reference module,
utils.py:404-411 over unet3d.py:190-259 ExtResNetBlock, :313-317 Encoder, :385-444 Decoder/Upsampling, :578 final_conv).

Forward (training mode) keeps every block's raw fp32 tensors and GroupNorm statistics on a tape; the normalised fp16
operands are recomputed here.  Per 'gcr' unit, with dy the gradient of its (post-ReLU) output:
    dz   = dy * [y > 0]                      -> fp16, per-tensor power-of-two scale          (semabs_unet_bwd_pack)
    dW   = sum_v dz[v] (x) GN(x)[v + tap]    -> split-K mma.sync reduction over padded voxels (semabs_conv3d_wgrad)
    dxn  = conv(dz, W adjoint)               -> the forward tcgen05 conv kernels              (semabs_conv3d[_halo])
    dx, dgamma, dbeta = GroupNorm backward   -> two bandwidth passes                          (semabs_groupnorm_bwd_*)
Max-pool / transposed-conv / final-conv gradients follow the same pattern (semabs_maxpool3d_2_bwd, conv kind 3 / 1).
Gradients are returned (not accumulated into .grad): torch autograd owns the accumulation.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .unet3d import F16, F32, F64, ExtResNetBlock, ResidualUNet3D, _pad16, _params_key, _split_pack

I32 = torch.int32
WGRAD_WS_BYTES = 256 << 20


class Tape:
    """Buffers of one training-mode forward that must survive until its backward. Pools are recycled between steps."""

    def __init__(self, pool: Dict):
        self.pool = pool
        self.blocks: Dict[str, dict] = {}
        self.meta: dict = {}
        self.extra: dict = {}

    def alloc(self, name, shape, dtype, dev):
        key = (name, tuple(int(s) for s in shape), dtype, str(dev))
        t = self.pool.get(key)
        if t is None:
            t = torch.empty(key[1], dtype=dtype, device=dev)
            self.pool[key] = t
        return t


class _Grad:
    """An fp32 gradient tensor with its (optional) device scale and |max| slot."""

    def __init__(self, t, scale=None, amax=None):
        self.t, self.scale, self.amax = t, scale, amax


class UNetBackward:
    def __init__(self, unet: ResidualUNet3D):
        self.unet = unet
        self._free_pools: List[Dict] = []
        self._bpk: Dict[str, torch.Tensor] = {}
        self._bpk_key = None
        self._slot_i = 0
        # data-gradient convolutions as dz_hi*w_hi + dz_lo*w_hi + dz_hi*w_lo (like the forward's precise mode): fp16
        # rounding of the adjoint WEIGHTS is a coherent error over all voxels, which single-pass operands turn into ~1e-2
        # relative errors of the (strongly cancelling) GroupNorm parameter gradients; the weight-gradient reduction only
        # sees per-voxel (incoherent) rounding and stays single pass
        # (follows the module's `precise` flag: precise=False = single fp16 operands everywhere, the counterpart of the
        # reference's --use_amp training mode, utils.py:78,291-294)
        self.precise = bool(unet.precise)

    # ---- tapes ------------------------------------------------------------------------------------------
    def new_tape(self) -> Tape:
        return Tape(self._free_pools.pop() if self._free_pools else {})

    def release(self, tape: Tape):
        if len(self._free_pools) < 4:
            self._free_pools.append(tape.pool)

    # ---- adjoint weight packs -----------------------------------------------------------------------------
    def _packed(self, dev):
        u = self.unet
        key = _params_key(u, dev, self.precise)
        if key == self._bpk_key:
            return self._bpk
        pk: Dict[str, torch.Tensor] = {}
        s = 2 if self.precise else 1

        def adj(w):  # conv.weight [Co,Ci,3,3,3] -> weight of the adjoint conv [Ci_pad, Co, 3,3,3] (taps flipped)
            co, ci = w.shape[:2]
            wa = torch.zeros(_pad16(ci), co, 3, 3, 3, device=dev)
            wa[:ci] = w.detach().to(dev, F32).flip(2, 3, 4).transpose(0, 1)
            return wa

        def block(prefix, blk: ExtResNetBlock):
            for j, sc in enumerate((blk.conv1, blk.conv2, blk.conv3), 1):
                wa = adj(sc.conv.weight)
                pk[f"{prefix}.wa{j}"] = _split_pack(wa.permute(0, 2, 3, 4, 1).reshape(wa.shape[0], -1), s)
                if wa.shape[0] in (16, 32) and wa.shape[1] in (16, 32):
                    pk[f"{prefix}.wah{j}"] = ops.pack_halo_weights(wa, s)

        for i, enc in enumerate(u.encoders):
            block(f"enc{i}", enc.basic_module)
        for i, dec in enumerate(u.decoders):
            block(f"dec{i}", dec.basic_module)
            w = dec.upsampling.upsample.weight  # [Ci, Co, 3,3,3]
            pk[f"dec{i}.up_wa"] = _split_pack(w.detach().to(dev, F32).permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1), s)
        fw = u.final_conv.weight  # [Co, Ci, 1,1,1]
        pk["final.wa"] = _split_pack(fw.detach().to(dev, F32).reshape(fw.shape[0], fw.shape[1]).t().contiguous(), s)
        pk["slot_k27"] = torch.arange(27, dtype=I32, device=dev)
        # transposed conv: launch A = taps with kx in {0,2} (x-parity 1 volume), launch B = taps with kx == 1
        pk["slot_kA"] = torch.tensor([(kz * 3 + ky) * 3 + kx for kz in range(3) for ky in range(3) for kx in (0, 2)], dtype=I32, device=dev)
        pk["slot_kB"] = torch.tensor([(kz * 3 + ky) * 3 + 1 for kz in range(3) for ky in range(3)], dtype=I32, device=dev)
        self._bpk, self._bpk_key = pk, key
        return pk

    # ---- scratch ----------------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype, dev, zero=False):
        u = self.unet
        key = ("bwd_" + name, tuple(int(s) for s in shape), dtype, str(dev))
        t = u._ws.get(key)
        if t is None:  # one buffer per (name, shape): the levels alternate, nothing is freed between them
            t = (torch.zeros if zero else torch.empty)(key[1], dtype=dtype, device=dev)
            u._ws[key] = t
        return t

    def _padded(self, name, N, dims, Cp, dev, volumes=1):
        """zero-initialised padded fp16 volume(s) with guard rows; returns (view starting at padded voxel 0, nvox rounded
        up to the kernel's K block, padded dims). Only interiors are ever written, so ring and guards stay zero."""
        PD, PH, PW = dims[0] + 2, dims[1] + 2, dims[2] + 2
        nv = volumes * N * PD * PH * PW
        guard = PH * PW + PW + 1
        KB = ops.WGRAD_KB
        nvr = (nv + KB - 1) // KB * KB
        total = guard + nvr + guard + KB + 2
        buf = self._buf(f"{name}_{N}_{PD}x{PH}x{PW}_{Cp}_{volumes}", (total, Cp), F16, dev, zero=True)
        return buf[guard:], nvr, (PD, PH, PW)

    def _slot(self):
        i = self._slot_i
        self._slot_i += 1
        assert i < self._amax.numel()
        return self._amax[i : i + 1], self._scales[i : i + 1]

    # ---- one 'gcr' unit -------------------------------------------------------------------------------------
    def _unit(self, pk, fpk, prefix, j, sc, g: _Grad, mask, x_raw, x_stats, *, N, dims, c_in_pad, c_in_real, c_out, dev,
              grads, dx, dx_accumulate=False, add: Optional[_Grad] = None, add_mask=None, dx_amax=None):
        u = self.unet
        D, H, W = dims
        S = D * H * W
        halo = W == 128 and c_in_pad in (16, 32) and c_out in (16, 32) and u.use_halo
        dz_pad, nvox, (PD, PH, PW) = self._padded("dzp", N, dims, c_out, dev)
        xn_pad, _, _ = self._padded("xnp", N, dims, c_in_pad, dev)
        sp = 2 if self.precise else 1
        dz_op = self._buf("dz_op", (N * S * sp * c_out,), F16, dev)
        _, s_dz = self._slot()
        ops.unet_bwd_pack(g.t, N=N, D=D, H=H, W=W, C=c_out, g_scale=g.scale, amax=g.amax, mask=mask, pad16=dz_pad, Cp=c_out,
                          op16=dz_op, op_layout=2 if halo else 1, op_splits=sp, scale_out=s_dz)
        gam, bet = fpk[f"{prefix}.g{j}"], fpk[f"{prefix}.b{j}"]
        groups = sc.num_groups
        ops.groupnorm_apply_padded(x_raw, x_stats, gam, bet, xn_pad, N=N, D=D, H=H, W=W, C=c_in_pad, C_real=c_in_real,
                                   groups=groups, Cp=c_in_pad)
        # weight gradient [Co, Ci, 3,3,3]
        gw = torch.empty_like(sc.conv.weight, device=dev, dtype=F32)
        seg_off, seg_sh, seg_slot = [], [], []
        for kd in range(3):
            for kh in range(3):
                seg_off.append((kd - 1) * PH * PW + (kh - 1) * PW - 1)
                seg_sh.append([0, 1, 2])
                seg_slot.append([(kd * 3 + kh) * 3 + kw for kw in range(3)])
        ops.conv3d_wgrad(dz_pad, xn_pad, lda=c_out, Ca=c_out, Ca_real=c_out, ldb=c_in_pad, Cb=c_in_pad, Cb_real=c_in_real,
                         nvox=nvox, seg_off=seg_off, seg_ntaps=[3] * 9, seg_sh=seg_sh, seg_slot=seg_slot, nslots=27,
                         slot_k=pk["slot_k27"], KT=27, workspace=self._wg_ws, scale=s_dz, grad=gw)
        grads[sc.conv.weight] = gw
        # data gradient through the conv: the forward kernels with the adjoint weights (single fp16 pass)
        dxn = self._buf("dxn", (N * S * c_in_pad,), F32, dev)
        common = dict(N=N, D=D, H=H, W=W, C_in=c_out, C_out=c_in_pad, a_splits=sp, w_splits=sp, precise=self.precise, out32=dxn)
        if halo:
            ops.conv3d_halo(dz_op, pk[f"{prefix}.wah{j}"], **common)
        else:
            ops.conv3d(dz_op, pk[f"{prefix}.wa{j}"], kind=ops.CONV_3X3X3, **common)
        # GroupNorm backward
        sums = self._buf("gn_sums", (N, c_in_pad, 2), F64, dev)
        sums.zero_()
        ops.groupnorm_bwd_reduce(dxn, x_raw, sums, N=N, S=S, C=c_in_pad)
        gn = sc.groupnorm
        dg, db = torch.empty_like(gn.weight, device=dev, dtype=F32), torch.empty_like(gn.bias, device=dev, dtype=F32)
        ops.groupnorm_param_grads(sums, N=N, S=S, C=c_in_pad, C_real=c_in_real, groups=groups, stats=x_stats, scale=s_dz,
                                  dgamma=dg, dbeta=db)
        grads[gn.weight], grads[gn.bias] = dg, db
        if dx is not None:
            ops.groupnorm_bwd_apply(dxn, x_raw, x_stats, gam, sums, dx, N=N, S=S, C=c_in_pad, C_real=c_in_real, groups=groups,
                                    dy_scale=s_dz, add=add.t if add is not None else None,
                                    add_scale=add.scale if add is not None else None, add_mask=add_mask,
                                    accumulate=dx_accumulate, amax=dx_amax)
        u.kernel_launches += 8

    def _block(self, pk, fpk, prefix, blk: ExtResNetBlock, rec, g_out: _Grad, *, N, dev, grads, dx, dx_accumulate, dx_amax):
        """ExtResNetBlock backward; writes (or accumulates) the gradient of the block input into dx (true scale)."""
        dims, c_in_pad, c_in_real, c_out = rec["dims"], rec["c_in_pad"], rec["c_in_real"], rec["c_out"]
        S = dims[0] * dims[1] * dims[2]
        st = rec["st"]
        kw = dict(N=N, dims=dims, dev=dev, grads=grads)
        # conv3 (no ReLU of its own: the mask is the block's final ReLU)
        d_o2 = self._buf(f"d_o2", (N * S * c_out,), F32, dev)
        a2, _ = self._slot()
        self._unit(pk, fpk, prefix, 3, blk.conv3, g_out, rec["out"], rec["o2"], st[1], c_in_pad=c_out, c_in_real=c_out,
                   c_out=c_out, dx=d_o2, dx_amax=a2, **kw)
        # conv2; the residual branch (out = relu(conv3 + o1)) joins the gradient of o1 here
        d_o1 = self._buf(f"d_o1", (N * S * c_out,), F32, dev)
        a1, _ = self._slot()
        self._unit(pk, fpk, prefix, 2, blk.conv2, _Grad(d_o2, None, a2), rec["o2"], rec["o1"], st[0], c_in_pad=c_out,
                   c_in_real=c_out, c_out=c_out, dx=d_o1, dx_amax=a1, add=g_out, add_mask=rec["out"], **kw)
        # conv1
        self._unit(pk, fpk, prefix, 1, blk.conv1, _Grad(d_o1, None, a1), rec["o1"], rec["x"], rec["x_stats"],
                   c_in_pad=c_in_pad, c_in_real=c_in_real, c_out=c_out, dx=dx, dx_accumulate=dx_accumulate, dx_amax=dx_amax,
                   **kw)

    # ---- whole network ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def backward(self, tape: Tape, d_out_cl: torch.Tensor, need_dx: bool = True, buckets=None):
        """d_out_cl: gradient of the channels-last output [N, S, out_channels] (true scale). Returns
        (dx channels-last [N, S, Cpad_in] or None, {parameter: gradient}).  `buckets` (train.GradientBuckets): the weight
        gradients of each resolution level are handed to it as soon as they are queued, so their all-reduce overlaps
        the levels still to come; the caller joins."""
        n_sent = 0

        def hand_over():
            nonlocal n_sent
            if buckets is not None:
                keys = list(grads)[n_sent:]
                buckets.submit(grads, keys)
                n_sent += len(keys)

        u = self.unet
        N, dims0, dev = tape.meta["N"], tape.meta["dims0"], tape.meta["dev"]
        pk, fpk = self._packed(dev), u._packed(dev)
        self._amax = self._buf("amax", (512,), I32, dev)
        self._scales = self._buf("scales", (512,), F32, dev)
        self._amax.zero_()
        self._slot_i = 0
        self._wg_ws = self._buf("wgrad_ws", (WGRAD_WS_BYTES // 4,), F32, dev)
        grads: Dict[torch.nn.Parameter, torch.Tensor] = {}
        L = len(u.encoders)
        f = u.f_maps

        # final 1x1x1 conv: y = W x + b, x = out of the last decoder block (or of the only encoder)
        last_prefix = f"dec{L - 2}" if L > 1 else "enc0"
        rec = tape.blocks[last_prefix]
        D, H, W = dims0
        S = D * H * W
        co, ci = u.out_channels, f[0]
        cop = _pad16(co)
        d_out_cl = d_out_cl.contiguous()
        a_y, s_y = self._slot()
        ops.absmax_f32(d_out_cl, a_y)
        dy_pad, nvox, _ = self._padded("dzp", N, dims0, cop, dev)
        x_pad, _, _ = self._padded("xnp", N, dims0, ci, dev)
        sp = 2 if self.precise else 1
        dy_op = self._buf("dz_op", (N * S * sp * co,), F16, dev)
        ops.unet_bwd_pack(d_out_cl, N=N, D=D, H=H, W=W, C=co, amax=a_y, pad16=dy_pad, Cp=cop, op16=dy_op, op_layout=1,
                          op_splits=sp, scale_out=s_y)
        ops.unet_bwd_pack(rec["out"], N=N, D=D, H=H, W=W, C=ci, pad16=x_pad, Cp=ci)
        gw = torch.empty_like(u.final_conv.weight, device=dev, dtype=F32)
        ops.conv3d_wgrad(dy_pad, x_pad, lda=cop, Ca=cop, Ca_real=co, ldb=ci, Cb=ci, Cb_real=ci, nvox=nvox, seg_off=[0],
                         seg_ntaps=[1], seg_sh=[[0, 0, 0]], seg_slot=[[0, 0, 0]], nslots=1, slot_k=pk["slot_k27"], KT=1,
                         workspace=self._wg_ws, scale=s_y, grad=gw)
        grads[u.final_conv.weight] = gw
        sums = self._buf("gn_sums", (N, co, 2), F64, dev)
        sums.zero_()
        ops.groupnorm_bwd_reduce(d_out_cl, None, sums, N=N, S=S, C=co)
        gb = torch.empty_like(u.final_conv.bias, device=dev, dtype=F32)
        ops.groupnorm_param_grads(sums, N=N, S=S, C=co, C_real=co, dbeta=gb)
        grads[u.final_conv.bias] = gb
        d_cur = self._buf("d_top", (N * S * ci,), F32, dev)
        ops.conv3d(dy_op, pk["final.wa"], kind=ops.CONV_1X1X1, N=N, D=D, H=H, W=W, C_in=co, C_out=ci, a_splits=sp, w_splits=sp,
                   precise=self.precise, out32=d_cur)
        a_c, _ = self._slot()
        ops.absmax_f32(d_cur, a_c)
        g_cur = _Grad(d_cur, s_y, a_c)
        u.kernel_launches += 7

        # decoders, top level first (reverse of the forward order)
        d_skip: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        for j in range(L - 2, -1, -1):
            dec = u.decoders[j]
            lvl = L - 2 - j
            rec = tape.blocks[f"dec{j}"]
            dims = rec["dims"]
            S = dims[0] * dims[1] * dims[2]
            c_out, c_in = f[lvl], f[lvl + 1]
            # gradient of `up` = conv-transpose output + skip: both consumers get it
            d_up = self._buf(f"d_skip{lvl}", (N * S * c_out,), F32, dev)
            a_up, s_up = self._slot()
            self._block(pk, fpk, f"dec{j}", dec.basic_module, rec, g_cur, N=N, dev=dev, grads=grads, dx=d_up,
                        dx_accumulate=False, dx_amax=a_up)
            d_skip[lvl] = (d_up, a_up)
            # transposed conv backward
            Dl, Hl, Wl = dims[0] // 2, dims[1] // 2, dims[2] // 2
            Sl = Dl * Hl * Wl
            low_prefix = f"dec{j - 1}" if j > 0 else f"enc{L - 1}"
            x_low = tape.blocks[low_prefix]["out"]
            dy_par, nvox_l, (PD, PH, PW) = self._padded("dyp", N, (Dl, Hl, Wl), c_out, dev, volumes=8)
            xl_pad, nvox_x, _ = self._padded("xlp", N, (Dl, Hl, Wl), c_in, dev)
            dy_op = self._buf("dz_op", (N * S * sp * c_out,), F16, dev)
            ops.unet_bwd_pack(d_up, N=N, D=dims[0], H=dims[1], W=dims[2], C=c_out, amax=a_up, pad16=dy_par, Cp=c_out, parity=True,
                              op16=dy_op, op_layout=1, op_splits=sp, scale_out=s_up)
            ops.unet_bwd_pack(x_low, N=N, D=Dl, H=Hl, W=Wl, C=c_in, pad16=xl_pad, Cp=c_in)
            sums = self._buf("gn_sums", (N, c_out, 2), F64, dev)
            sums.zero_()
            ops.groupnorm_bwd_reduce(d_up, None, sums, N=N, S=S, C=c_out)
            up = dec.upsampling.upsample
            gb = torch.empty_like(up.bias, device=dev, dtype=F32)
            ops.groupnorm_param_grads(sums, N=N, S=S, C=c_out, C_real=c_out, dbeta=gb)
            grads[up.bias] = gb
            # weight [Ci, Co, 3,3,3]: dW[ci,co,k] = sum_i x[i,ci] dy[2i-1+k, co]; k=0 -> parity 1, j=i-1; k=1 -> parity 0;
            # k=2 -> parity 1, j=i.  Taps sharing an x-parity volume form one segment.
            gw = torch.empty_like(up.weight, device=dev, dtype=F32)
            vol = N * PD * PH * PW
            par = lambda k: 0 if k == 1 else 1
            sh = lambda k: -1 if k == 0 else 0
            offA, offB = [], []
            for kz in range(3):
                for ky in range(3):
                    base = sh(kz) * PH * PW + sh(ky) * PW
                    offA.append(((par(kz) << 2) | (par(ky) << 1) | 1) * vol + base - 1)  # kx=0 -> shift 0, kx=2 -> shift 1
                    offB.append(((par(kz) << 2) | (par(ky) << 1) | 0) * vol + base)
            wk = dict(lda=c_in, Ca=c_in, Ca_real=c_in, ldb=c_out, Cb=c_out, Cb_real=c_out, nvox=nvox_x, KT=27,
                      workspace=self._wg_ws, scale=s_up, grad=gw)
            ops.conv3d_wgrad(xl_pad, dy_par, seg_off=offA, seg_ntaps=[2] * 9, seg_sh=[[0, 1, 0]] * 9,
                             seg_slot=[[2 * s, 2 * s + 1, 0] for s in range(9)], nslots=18, slot_k=pk["slot_kA"], **wk)
            ops.conv3d_wgrad(xl_pad, dy_par, seg_off=offB, seg_ntaps=[1] * 9, seg_sh=[[0, 0, 0]] * 9,
                             seg_slot=[[s, 0, 0] for s in range(9)], nslots=9, slot_k=pk["slot_kB"], **wk)
            grads[up.weight] = gw
            # data gradient: stride-2 conv over d_up, one input-parity class per launch, chained through the residual
            d_low = self._buf(f"d_low{lvl + 1}", (N * Sl * c_in,), F32, dev)
            for q in range(8):
                ops.conv3d(dy_op, pk[f"dec{j}.up_wa"], kind=ops.CONV_TRANSPOSE_ADJOINT, parity=q, N=N, D=Dl, H=Hl, W=Wl,
                           C_in=c_out, C_out=c_in, a_splits=sp, w_splits=sp, precise=self.precise, out32=d_low,
                           residual=d_low if q > 0 else None)
            a_l, _ = self._slot()
            ops.absmax_f32(d_low, a_l)
            g_cur = _Grad(d_low, s_up, a_l)
            u.kernel_launches += 16
            hand_over()

        # encoders, deepest first
        dx_in = None
        for i in range(L - 1, -1, -1):
            enc = u.encoders[i]
            rec = tape.blocks[f"enc{i}"]
            dims = rec["dims"]
            S = dims[0] * dims[1] * dims[2]
            if i < L - 1:
                t, a = d_skip[i]
                g_cur = _Grad(t, None, a)  # skip gradient + (already accumulated) pooled-path gradient
            if L == 1:
                pass  # g_cur is the final conv's data gradient
            c_in_pad = rec["c_in_pad"]
            if i > 0 or need_dx:
                d_in = self._buf("d_in", (N * S * c_in_pad,), F32, dev) if i > 0 else torch.empty(N, S, c_in_pad, device=dev)
            else:
                d_in = None
            self._block(pk, fpk, f"enc{i}", enc.basic_module, rec, g_cur, N=N, dev=dev, grads=grads, dx=d_in,
                        dx_accumulate=False, dx_amax=None)
            if i > 0:
                # MaxPool3d(2) backward into the gradient of the previous encoder's output (joins its skip gradient)
                t, a = d_skip[i - 1]
                Dp, Hp, Wp = dims[0] * 2, dims[1] * 2, dims[2] * 2
                ops.maxpool3d_2_bwd(d_in, tape.blocks[f"enc{i - 1}"]["out"], t, N=N, D=Dp, H=Hp, W=Wp, C=c_in_pad,
                                    accumulate=True, amax=a)
                u.kernel_launches += 1
            else:
                dx_in = d_in
            hand_over()
        return dx_in, grads


class _UNetFn(torch.autograd.Function):
    """autograd node: ResidualUNet3D.forward (NCDHW in/out) with the CUDA forward / backward above."""

    @staticmethod
    def forward(ctx, x, unet, *params):
        N, C, D, H, W = x.shape
        dev = x.device
        S = D * H * W
        bw = unet._bwd()
        tape = bw.new_tape()
        cpad = _pad16(C)
        g_in = unet.encoders[0].basic_module.conv1.num_groups
        raw = tape.alloc("l0_in", (N, S, cpad), F32, dev)
        st = tape.alloc("l0_pst", (N, 8, 2), F64, dev)
        st.zero_()
        ops.ncdhw_to_ndhwc(x.contiguous().float(), raw, N=N, S=S, C=C, Cpad=cpad, groups=g_in, stats=st)
        out_cl = unet.forward_channels_last(raw, st, N, (D, H, W), dev, tape=tape)
        y = torch.empty(N, unet.out_channels, D, H, W, device=dev)
        ops.ndhwc_to_ncdhw(out_cl, y, N=N, S=S, C=unet.out_channels)
        ctx.unet, ctx.tape, ctx.shape = unet, tape, (N, C, D, H, W)
        ctx.params = params
        return y

    @staticmethod
    def backward(ctx, dy):
        unet, tape = ctx.unet, ctx.tape
        N, C, D, H, W = ctx.shape
        S = D * H * W
        dev = dy.device
        co = unet.out_channels
        dy_cl = torch.empty(N, S, co, device=dev)
        ops.ncdhw_to_ndhwc(dy.contiguous().float(), dy_cl, N=N, S=S, C=co, Cpad=co)
        bw = unet._bwd()
        need_dx = ctx.needs_input_grad[0]
        from .train import GradientBuckets

        buckets = GradientBuckets.active()
        dx_cl, grads = bw.backward(tape, dy_cl, need_dx=need_dx, buckets=buckets)
        bw.release(tape)
        dx = None
        if need_dx:
            cpad = _pad16(C)
            if cpad == C:
                dx = torch.empty(N, C, D, H, W, device=dev)
                ops.ndhwc_to_ncdhw(dx_cl, dx, N=N, S=S, C=C)
            else:
                full = torch.empty(N, cpad, D, H, W, device=dev)
                ops.ndhwc_to_ncdhw(dx_cl, full, N=N, S=S, C=cpad)
                dx = full[:, :C].contiguous()
        if buckets is not None:
            buckets.join(grads)
        return (dx, None) + tuple(grads.get(p) for p in ctx.params)
