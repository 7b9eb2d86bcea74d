"""Device-side CLIP engine: stream-ordered sequence of C-ABI kernel calls that replaces
ClipGradcam.forward + ClipGradcam.interpret (reference: CLIP/clip/clip_gradcam.py:58-132) and CLIP.encode_text.

What is different from the reference on purpose (verified against the oracle):
  * ONE hand-written backward sweep from the last block down to block 11 for all P label cotangents batched in
    the GEMM M dimension (the reference calls torch.autograd.grad P x 13 times, re-traversing later blocks);
  * the rollout keeps only row 0 of R (the only row read at clip_gradcam.py:127): r <- r + r·cam, applied in
    the same top-down order as the backward sweep, so no [T,T] matrix product is ever formed.
torch is used for allocation only.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .. import ops
from .model import BlockWeights, ClipDeviceWeights, TowerWeights

F16, F32 = torch.float16, torch.float32


class _Workspace:
    """Shape-keyed cache of device buffers so that steady-state calls allocate nothing."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[Tuple, torch.Tensor] = {}

    def get(self, name, shape, dtype=F32):
        key = (name, tuple(int(s) for s in shape), dtype)
        t = self.bufs.get(key)
        if t is None:
            # drop stale buffers of the same name (different batch shape) to bound memory
            for k in [k for k in self.bufs if k[0] == name]:
                del self.bufs[k]
            t = torch.empty(key[1], dtype=dtype, device=self.device)
            self.bufs[key] = t
        return t


class ClipEngine:
    def __init__(self, weights: ClipDeviceWeights, device, num_layers: int = 10, fwd_splits: int = 2, bwd_splits: int = 1):
        self.w = weights
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("semabs_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        self.start_block = num_layers + 1  # `if i <= self.num_layers: continue` (clip_gradcam.py:86-87)
        self.fwd_splits = fwd_splits
        self.bwd_splits = bwd_splits
        self.ws = _Workspace(self.device)
        self.kernel_launches = 0
        self.sparse_last_block = True
        self.tc_attention = True  # tcgen05 attention forward (T <= 272); False = the mma.sync kernel of vit_attn.cu
        self.tc_attention_bwd = True  # tcgen05 attention backward (T <= 272); False = the mma.sync kernels
        # dense-block LayerNorm cotangents as fp16 (one more fp16 rounding of the class every dgrad operand already has with
        # bwd_splits = 1; with bwd_splits = 2 — the fp32-grade backward — they stay fp32)
        self.ln_bwd_f16 = bwd_splits == 1

    # ------------------------------------------------------------------------------------------------
    def _block_forward(self, blk: BlockWeights, x, x_next, *, n_seq, T, H, d, causal, saved: Optional[dict]):
        s = self.fwd_splits
        M = n_seq * T
        ws = self.ws
        h16 = ws.get("h16", (M, s * d), F16)
        mean1 = rstd1 = mean2 = rstd2 = probs16 = qkv16 = o32 = gg16 = None
        if saved is not None:
            mean1, rstd1 = saved["mean1"], saved["rstd1"]
            mean2, rstd2 = saved["mean2"], saved["rstd2"]
            probs16, qkv16, o32, gg16 = saved["probs16"], saved["qkv16"], saved["o32"], saved["gelu_grad16"]
            x_mid = saved["x_mid"]
        else:
            x_mid = ws.get("x_mid", (M, d))
        ops.layernorm_fwd(x, blk.ln1_g, blk.ln1_b, M=M, d=d, y16=h16, mean=mean1, rstd=rstd1, splits=s)
        o16 = ws.get("o16", (M, s * d), F16)
        if qkv16 is None:
            qkv16 = ws.get("qkv16_scratch", (M, s * 3 * d), F16)
        if self.tc_attention and T <= 272:
            # tcgen05 attention: q,k,v leave the GEMM as fp16 hi|lo rows (hi*hi + lo*hi + hi*lo = fp32-grade products);
            # the hi part doubles as the MMA operand of the attention backward
            ops.gemm_f16(h16, blk.w_in, a_splits=s, bias=blk.b_in, out_f16=qkv16, out_f16_splits=s, scale_cols=d, scale=0.125)
            ops.attn_fwd_tc(qkv16, in_splits=s, B=n_seq, T=T, H=H, probs16=probs16, o32=o32, o16=o16, o_splits=s,
                            causal=causal)
        else:
            # fp32 q,k,v feed the mma.sync forward; the fp16 copy is the MMA operand of the attention backward
            qkv = ws.get("qkv", (M, 3 * d))
            ops.gemm_f16(h16, blk.w_in, a_splits=s, bias=blk.b_in, out_f32=qkv, out_f16=qkv16[:, : 3 * d], scale_cols=d,
                         scale=0.125)
            ops.attn_fwd(qkv, B=n_seq, T=T, H=H, probs16=probs16, o32=o32, o16=o16, causal=causal, splits=s)
        ops.gemm_f16(o16, blk.w_out, a_splits=s, bias=blk.b_out, residual=x, out_f32=x_mid)
        ops.layernorm_fwd(x_mid, blk.ln2_g, blk.ln2_b, M=M, d=d, y16=h16, mean=mean2, rstd=rstd2, splits=s)
        g16 = ws.get("g16", (M, s * 4 * d), F16)
        # the backward sweep needs d quickgelu / du only: keep it as fp16 instead of the fp32 pre-activation
        ops.gemm_f16(h16, blk.w_fc, a_splits=s, bias=blk.b_fc, act=ops.ACT_QUICKGELU, out_f16=g16, out_f16_splits=s,
                     out_aux16=gg16)
        ops.gemm_f16(g16, blk.w_proj, a_splits=s, bias=blk.b_proj, residual=x_mid, out_f32=x_next)
        self.kernel_launches += 7

    # ------------------------------------------------------------------------------------------------
    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens [n, 77] integer -> features [n, E] fp32 (CLIP.encode_text, model_explainability.py:468-482)."""
        w, ws, s = self.w, self.ws, self.fwd_splits
        tw: TowerWeights = w.text
        n, ctx = tokens.shape
        d, H = tw.width, tw.heads
        tok = tokens.to(device=self.device, dtype=torch.int32).contiguous()
        xa = ws.get("tx_a", (n * ctx, d))
        xb = ws.get("tx_b", (n * ctx, d))
        ops.text_embed(tok, w.token_embedding, w.text_pos, xa, n, ctx, d)
        x, x_next = xa, xb
        for blk in tw.blocks:
            self._block_forward(blk, x, x_next, n_seq=n, T=ctx, H=H, d=d, causal=True, saved=None)
            x, x_next = x_next, x
        # features of the EOT token = highest token id in each sequence (model_explainability.py:478-480)
        eot = tok.argmax(dim=-1)
        rows = x.view(n, ctx, d)[torch.arange(n, device=self.device), eot].contiguous()
        y16 = ws.get("t_y16", (n, s * d), F16)
        ops.layernorm_fwd(rows, w.ln_final_g, w.ln_final_b, M=n, d=d, y16=y16, splits=s)
        feat = torch.empty(n, w.embed_dim, device=self.device)
        ops.gemm_f16(y16, w.text_projT, a_splits=s, out_f32=feat)
        self.kernel_launches += 3
        return feat

    def zeroshot_weights(self, tokens: torch.Tensor, n_classes: int, n_templates: int) -> torch.Tensor:
        """[n_classes*n_templates, 77] tokens -> W [E, n_classes] (zeroshot_classifier, clip_gradcam.py:12-27)."""
        feat = self.encode_text(tokens)
        W = torch.empty(self.w.embed_dim, n_classes, device=self.device)
        ops.zeroshot_weights(feat, W, n_classes, n_templates, self.w.embed_dim)
        self.kernel_launches += 1
        return W

    # ------------------------------------------------------------------------------------------------
    def _saved(self, i, B, T, d, H):
        ws = self.ws
        M = B * T
        return {
            "x_in": None,
            "mean1": ws.get(f"s{i}_mean1", (M,)), "rstd1": ws.get(f"s{i}_rstd1", (M,)),
            "mean2": ws.get(f"s{i}_mean2", (M,)), "rstd2": ws.get(f"s{i}_rstd2", (M,)),
            "qkv16": ws.get(f"s{i}_qkv16", (M, self.fwd_splits * 3 * d), F16),
            "probs16": ws.get(f"s{i}_probs16", (B * H, T, (T + 15) // 16 * 16), F16),
            "o32": ws.get(f"s{i}_o32", (M, d)), "x_mid": ws.get(f"s{i}_xmid", (M, d)),
            "gelu_grad16": ws.get(f"s{i}_gg16", (M, 4 * d), F16),
        }  # fmt: skip

    def encode_image(self, tiles: torch.Tensor, keep_for_backward: bool = False):
        """tiles [B,3,R,R] fp32 on device -> (features [B,E], state). VisionTransformer.forward,
        model_explainability.py:324-355."""
        w, ws, s = self.w, self.ws, self.fwd_splits
        vt: TowerWeights = w.visual
        B = tiles.shape[0]
        assert tiles.shape[1] == 3 and tiles.shape[2] == tiles.shape[3] == w.input_resolution, tiles.shape
        d, H, T, g = vt.width, vt.heads, w.tokens, w.grid
        M = B * T
        L = len(vt.blocks)
        a_pe = ws.get("im2col", (B * g * g, s * w.kpad), F16)
        ops.vit_im2col(tiles, a_pe, w.patch, w.kpad, s)
        pe = ws.get("pe", (B * g * g, d))
        ops.gemm_f16(a_pe, w.conv_w, a_splits=s, out_f32=pe)
        first_keep = self.start_block if keep_for_backward else L + 1
        saved = {}

        def x_buffer(i):
            # input buffer of block i: kept when block i's LN1 backward needs it
            if i > first_keep and i < L:
                return ws.get(f"s{i}_xin", (M, d))
            return ws.get(f"x_pp{i % 2}", (M, d))

        x = x_buffer(0)
        ops.vit_embed_lnpre(pe, w.cls, w.pos, w.ln_pre_g, w.ln_pre_b, x, B, T, d)
        self.kernel_launches += 3
        for i, blk in enumerate(vt.blocks):
            sv = None
            if i >= first_keep:
                sv = self._saved(i, B, T, d, H)
                sv["x_in"] = x
                saved[i] = sv
            x_next = x_buffer(i + 1) if i + 1 < L else ws.get("x_final", (M, d))
            self._block_forward(blk, x, x_next, n_seq=B, T=T, H=H, d=d, causal=False, saved=sv)
            x = x_next
        cls16 = ws.get("cls16", (B, s * d), F16)
        mean_p, rstd_p = ws.get("mean_p", (B,)), ws.get("rstd_p", (B,))
        ops.layernorm_fwd(x, w.ln_post_g, w.ln_post_b, M=B, d=d, x_stride=T * d, y16=cls16, mean=mean_p, rstd=rstd_p,
                          splits=s)
        f = ws.get("img_feat", (B, w.embed_dim))
        ops.gemm_f16(cls16, w.projT, a_splits=s, out_f32=f)
        self.kernel_launches += 2
        state = {"saved": saved, "x_final": x, "mean_p": mean_p, "rstd_p": rstd_p, "f": f, "B": B}
        return f, state

    # ------------------------------------------------------------------------------------------------
    def relevancy(self, tiles: torch.Tensor, W: torch.Tensor, positive_attn_only: bool = True,
                  return_logits: bool = False):
        """ClipGradcam.forward + interpret: tiles [B,3,R,R], W [E,P] -> relevance [P,B,g,g] fp32 (device)."""
        w, ws, sb = self.w, self.ws, self.bwd_splits
        vt = w.visual
        d, H, T, g, E = vt.width, vt.heads, w.tokens, w.grid, w.embed_dim
        L = len(vt.blocks)
        assert W.dtype == F32 and W.shape[0] == E and W.is_contiguous()
        P = W.shape[1]
        f, st = self.encode_image(tiles, keep_for_backward=True)
        B = st["B"]
        PB, Mb = P * B, P * B * T

        logits = ws.get("logits", (B, P))
        seed16 = ws.get("seed16", (PB, sb * E), F16)
        ops.clip_logit_seed(f, W, B=B, P=P, E=E, logits=logits, seed16=seed16, splits=sb)
        dy = ws.get("dy_cls", (PB, d))
        ops.gemm_f16(seed16, w.proj, a_splits=sb, out_f32=dy)
        dx = ws.get("dx_a", (Mb, d))
        dx16 = ws.get("dx16_a", (Mb, sb * d), F16)
        r = ws.get("rollout_r", (PB, T))
        ops.rollout_init(r, PB, T)
        self.kernel_launches += 6

        du16 = ws.get("du16", (Mb, sb * 4 * d), F16)
        # cotangent of the LayerNorm outputs: fp16 straight from the dgrad GEMM epilogue (the LayerNorm backward is HBM-bound)
        dh = ws.get("dh16", (Mb, d), F16) if self.ln_bwd_f16 else ws.get("dh", (Mb, d))
        dh_out = dict(out_f16=dh) if self.ln_bwd_f16 else dict(out_f32=dh)
        dxm = ws.get("dx_mid", (Mb, d))
        dxm16 = ws.get("dx_mid16", (Mb, sb * d), F16)
        dO16 = ws.get("dO16", (Mb, d), F16)
        delta = ws.get("attn_delta", (PB * H, T))
        wpart = ws.get("attn_wpart", (PB * H, T))
        dqkv16 = ws.get("dqkv16", (Mb, sb * 3 * d), F16) if L - 1 > self.start_block else None

        # ---- last block: its output cotangent is non-zero at the class token only (the logits read x[:,0]), so the
        # MLP / out-proj dgrads and both LayerNorm backwards run on P*B rows instead of P*B*T, and the attention
        # backward is vector work (semabs_attn_bwd_cls). From the block below everything is dense.
        first = L - 1
        if self.sparse_last_block:
            blk, sv = vt.blocks[first], st["saved"][first]
            need = first > self.start_block
            dxc = ws.get("cls_dx", (PB, d))
            dxc16 = ws.get("cls_dx16", (PB, sb * d), F16)
            ops.layernorm_bwd(dy, st["x_final"], st["mean_p"], st["rstd_p"], w.ln_post_g, dxc, M=PB, d=d, x_rows=B,
                              x_stride=T * d, dx16=dxc16, splits=sb)
            gg_cls = sv["gelu_grad16"].view(B, T, 4 * d)[:, 0, :]  # strided view: class-token rows
            duc16 = ws.get("cls_du16", (PB, sb * 4 * d), F16)
            ops.gemm_f16(dxc16, blk.w_projT, a_splits=sb, aux16=gg_cls, act=ops.ACT_MUL_AUX16, out_f16=duc16,
                         out_f16_splits=sb)
            dhc = ws.get("cls_dh", (PB, d))
            ops.gemm_f16(duc16, blk.w_fcT, a_splits=sb, out_f32=dhc)
            mean2c = sv["mean2"].view(B, T)[:, 0].contiguous()
            rstd2c = sv["rstd2"].view(B, T)[:, 0].contiguous()
            dxmc = ws.get("cls_dxm", (PB, d))
            dxmc16 = ws.get("cls_dxm16", (PB, sb * d), F16)
            ops.layernorm_bwd(dhc, sv["x_mid"], mean2c, rstd2c, blk.ln2_g, dxmc, M=PB, d=d, x_rows=B, x_stride=T * d,
                              dres=dxc, dx16=dxmc16, splits=sb)
            dOc16 = ws.get("cls_dO16", (PB, d), F16)
            ops.gemm_f16(dxmc16, blk.w_outT, a_splits=sb, out_f16=dOc16)
            ops.attn_bwd_cls(sv["qkv16"], sv["probs16"], dOc16, d, r, wpart, dqkv16 if need else None, P=P, B=B, T=T, H=H,
                             splits=sb, positive_only=positive_attn_only, need_dqkv=need)
            ops.rollout_update(r, wpart, PB, H, T)
            self.kernel_launches += 8
            if need:
                dxm.zero_()
                dxm.view(PB, T, d)[:, 0, :].copy_(dxmc)  # scatter of the class-token rows (data movement)
                ops.gemm_f16(dqkv16, blk.w_inT, a_splits=sb, **dh_out)
                ops.layernorm_bwd(dh, sv["x_in"], sv["mean1"], sv["rstd1"], blk.ln1_g, dx, M=Mb, d=d, x_rows=B * T,
                                  dres=dxm, dx16=dx16, splits=sb)
                self.kernel_launches += 2
            first -= 1
        else:
            dx.zero_()
            dx16.zero_()
            ops.layernorm_bwd(dy, st["x_final"], st["mean_p"], st["rstd_p"], w.ln_post_g, dx, M=PB, d=d, x_rows=B,
                              x_stride=T * d, out_stride=T * d, dx16=dx16, out16_stride=T * sb * d, splits=sb)

        for i in range(first, self.start_block - 1, -1):
            blk, sv = vt.blocks[i], st["saved"][i]
            # x_out = x_mid + c_proj(quickgelu(c_fc(ln_2(x_mid))))
            ops.gemm_f16(dx16, blk.w_projT, a_splits=sb, aux16=sv["gelu_grad16"], act=ops.ACT_MUL_AUX16, out_f16=du16,
                         out_f16_splits=sb)
            ops.gemm_f16(du16, blk.w_fcT, a_splits=sb, **dh_out)
            ops.layernorm_bwd(dh, sv["x_mid"], sv["mean2"], sv["rstd2"], blk.ln2_g, dxm, M=Mb, d=d, x_rows=B * T,
                              dres=dx, dx16=dxm16, splits=sb)
            # x_mid = x_in + out_proj(attn(ln_1(x_in)))
            ops.gemm_f16(dxm16, blk.w_outT, a_splits=sb, out_f16=dO16)
            need = i > self.start_block
            attn_bwd = ops.attn_bwd_tc if (self.tc_attention_bwd and T <= 272) else ops.attn_bwd
            attn_bwd(sv["qkv16"], sv["probs16"], sv["o32"], dO16, d, r, delta, wpart, dqkv16 if need else None, P=P,
                     B=B, T=T, H=H, splits=sb, positive_only=positive_attn_only, need_dqkv=need)
            ops.rollout_update(r, wpart, PB, H, T)
            self.kernel_launches += 8 if need else 7
            if need:
                ops.gemm_f16(dqkv16, blk.w_inT, a_splits=sb, **dh_out)
                ops.layernorm_bwd(dh, sv["x_in"], sv["mean1"], sv["rstd1"], blk.ln1_g, dx, M=Mb, d=d, x_rows=B * T,
                                  dres=dxm, dx16=dx16, splits=sb)
                self.kernel_launches += 2
        rel = r.view(P, B, T)[:, :, 1:].reshape(P, B, g, g)
        if return_logits:
            return rel, logits
        return rel
