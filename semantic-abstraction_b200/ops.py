"""Thin typed wrappers: torch device tensors in, C-ABI calls out. One function per entry point."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_MUL_AUX16, ACT_NONE, ACT_QUICKGELU, CALL_PROFILE, GemmEpilogue, check, f32, i32, lib, ptr, stream_ptr


class _GemmProfile:
    """Optional CUDA-event bracket around every GEMM launch (bench.py's live roofline measurement)."""

    def __init__(self):
        self.on = False
        self.recs = []

    def enable(self):
        self.on, self.recs = True, []

    def collect(self):
        torch.cuda.synchronize()
        ms = sum(e0.elapsed_time(e1) for e0, e1, _ in self.recs)
        flops = sum(f for _, _, f in self.recs)
        n = len(self.recs)
        self.on, self.recs = False, []
        return {"ms": ms, "launches": n, "tflops": (flops / (ms * 1e-3) / 1e12) if ms > 0 else 0.0, "flops": flops}


GEMM_PROFILE = _GemmProfile()


def gemm_f16(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    a_splits: int = 1,
    bias: torch.Tensor | None = None,
    residual: torch.Tensor | None = None,
    aux16: torch.Tensor | None = None,
    out_aux16: torch.Tensor | None = None,
    act: int = ACT_NONE,
    out_f32: torch.Tensor | None = None,
    out_f16: torch.Tensor | None = None,
    out_f16_splits: int = 1,
    scale_cols: int = 0,
    scale: float = 1.0,
):
    """C[M,N] = epi(A[M, a_splits*K] @ B[N,K]^T); see semabs_gemm_f16 in include/semabs_b200.h."""
    assert a.dtype == torch.float16 and b.dtype == torch.float16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    M = a.shape[0]
    N, K = b.shape
    assert a.shape[1] == a_splits * K, (a.shape, b.shape, a_splits)
    ep = GemmEpilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    ep.residual = residual.data_ptr() if residual is not None else None
    ep.aux16 = aux16.data_ptr() if aux16 is not None else None
    ep.aux_rows = aux16.shape[0] if aux16 is not None else 0
    ep.ld_aux = aux16.stride(0) if aux16 is not None else 0
    ep.out_aux16 = out_aux16.data_ptr() if out_aux16 is not None else None
    ep.ld_out_aux = out_aux16.stride(0) if out_aux16 is not None else 0
    assert aux16 is None or aux16.dtype == torch.float16
    assert out_aux16 is None or (out_aux16.dtype == torch.float16 and out_aux16.shape == (M, N))
    ep.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
    ep.ld_out = out_f32.stride(0) if out_f32 is not None else (residual.stride(0) if residual is not None else 0)
    ep.out_f16 = out_f16.data_ptr() if out_f16 is not None else None
    ep.ld_out16 = out_f16.stride(0) if out_f16 is not None else 0
    ep.out_f16_splits = out_f16_splits
    ep.act = act
    ep.scale_cols = scale_cols
    ep.scale = scale
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (M, N)
        assert out_f32 is None or residual.stride(0) == out_f32.stride(0)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.shape == (M, N)
    if out_f16 is not None:
        assert out_f16.dtype == torch.float16 and out_f16.shape == (M, out_f16_splits * N)
    CALL_PROFILE.note("semabs_gemm_f16", flops=2.0 * M * N * K)
    prof = GEMM_PROFILE.on
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(
        lib().semabs_gemm_f16(
            ptr(a), i32(a.stride(0)), ptr(b), i32(b.stride(0)), i32(M), i32(N), i32(K), i32(a_splits),
            C.byref(ep), stream_ptr(),
        )
    )
    if prof:
        e1.record()
        GEMM_PROFILE.recs.append((e0, e1, 2.0 * M * N * K))  # algorithmic FLOPs (hi/lo passes not double counted)
    return out_f32 if out_f32 is not None else out_f16


def set_gemm_pair(enable: bool) -> None:
    """CTA-pair (cta_group::2) GEMM kernel for the large shapes on / off (default on); see semabs_set_gemm_pair."""
    check(lib().semabs_set_gemm_pair(i32(int(enable))))


def split_f16(x: torch.Tensor) -> torch.Tensor:
    """[M,K] fp32 -> [M,2K] fp16 (hi | lo). Test helper; product kernels emit the split in their epilogues."""
    hi = x.half()
    lo = (x - hi.float()).half()
    return torch.cat([hi, lo], dim=1).contiguous()


# ---------------------------------------------------------------------------------------------------------
# ViT / text transformer stages
# ---------------------------------------------------------------------------------------------------------
def _i64(v):
    return C.c_int64(int(v))


def vit_im2col(tiles, out16, patch, kpad, splits):
    B, _, R, _ = tiles.shape
    assert tiles.dtype == torch.float32 and tiles.is_contiguous()
    check(lib().semabs_vit_im2col(ptr(tiles), ptr(out16), i32(B), i32(R), i32(patch), i32(kpad), i32(splits), stream_ptr()))


def layernorm_fwd(x, gamma, beta, *, M, d, x_stride=None, y32=None, y16=None, mean=None, rstd=None, splits=1):
    CALL_PROFILE.note("semabs_layernorm_fwd", bytes=M * d * (4 + (2 * splits if y16 is not None else 0) + (4 if y32 is not None else 0)))
    check(
        lib().semabs_layernorm_fwd(
            ptr(x), _i64(d if x_stride is None else x_stride), ptr(gamma), ptr(beta), ptr(y32), ptr(y16), ptr(mean),
            ptr(rstd), i32(M), i32(d), i32(splits), stream_ptr(),
        )
    )


def vit_embed_lnpre(patch, cls, pos, gamma, beta, x_out, B, T, d):
    check(lib().semabs_vit_embed_lnpre(ptr(patch), ptr(cls), ptr(pos), ptr(gamma), ptr(beta), ptr(x_out), i32(B), i32(T), i32(d), stream_ptr()))


def layernorm_bwd(dy, x, mean, rstd, gamma, dx32, *, M, d, x_rows, x_stride=None, dres=None, out_stride=None, dx16=None,
                  out16_stride=None, splits=1):
    """dy: fp32 [M, d] or fp16 [M, d] (semabs_layernorm_bwd_h)."""
    half = dy.dtype == torch.float16
    CALL_PROFILE.note("semabs_layernorm_bwd_h" if half else "semabs_layernorm_bwd",
                      bytes=M * d * ((2 if half else 4) + 4 + (4 if dres is not None else 0) + (2 * splits if dx16 is not None else 0))
                      + x_rows * d * 4)
    fn = lib().semabs_layernorm_bwd_h if half else lib().semabs_layernorm_bwd
    check(
        fn(
            ptr(dy), ptr(dres), ptr(x), _i64(d if x_stride is None else x_stride), i32(x_rows), ptr(mean), ptr(rstd),
            ptr(gamma), ptr(dx32), _i64(d if out_stride is None else out_stride), ptr(dx16),
            _i64(splits * d if out16_stride is None else out16_stride), i32(M), i32(d), i32(splits), stream_ptr(),
        )
    )


def attn_fwd(qkv, *, B, T, H, probs=None, probs16=None, o32=None, o16=None, causal=False, splits=1):
    ldp = probs16.shape[-1] if probs16 is not None else 0
    check(lib().semabs_attn_fwd(ptr(qkv), ptr(probs), ptr(probs16), i32(ldp), ptr(o32), ptr(o16), i32(B), i32(T), i32(H),
                                i32(int(causal)), i32(splits), stream_ptr()))


def attn_fwd_tc(qkv16, *, in_splits, B, T, H, probs16=None, o32=None, o16=None, o_splits=1, causal=False):
    """tcgen05 attention forward; qkv16 [B*T, in_splits*3d] fp16 rows [hi | lo]."""
    CALL_PROFILE.note("semabs_attn_fwd_tc", flops=4.0 * T * T * 64 * H * B)  # S = QK^T and O = PV
    assert qkv16.dtype == torch.float16 and qkv16.is_contiguous()
    ldp = probs16.shape[-1] if probs16 is not None else 0
    check(lib().semabs_attn_fwd_tc(ptr(qkv16), i32(in_splits), ptr(probs16), i32(ldp), ptr(o32), ptr(o16), i32(o_splits),
                                   i32(B), i32(T), i32(H), i32(int(causal)), stream_ptr()))


def selftest_ts_mma(A16, B16, D, lbo=16, sbo=1024):
    check(lib().semabs_selftest_ts_mma(ptr(A16), ptr(B16), ptr(D), i32(A16.shape[1]), i32(lbo), i32(sbo), stream_ptr()))


def attn_bwd(qkv16, probs16, o32, dO16, ld_do, r, delta_ws, wpart, dqkv16, *, P, B, T, H, splits=1, positive_only=True,
             need_dqkv=True):
    assert qkv16.dtype == torch.float16 and probs16.dtype == torch.float16
    check(
        lib().semabs_attn_bwd(
            ptr(qkv16), i32(qkv16.stride(0)), ptr(probs16), i32(probs16.shape[-1]), ptr(o32), ptr(dO16), i32(ld_do), ptr(r), ptr(delta_ws),
            ptr(wpart), ptr(dqkv16), i32(P), i32(B), i32(T), i32(H), i32(splits), i32(int(positive_only)),
            i32(int(need_dqkv)), stream_ptr(),
        )
    )


ATTN_BWD_GENERATION = 3  # 3 = vit_attn_bwd3.cu (product path: chunk-pipelined passes); 2 = vit_attn_bwd2.cu; 1 = the first tcgen05
# kernels (cross-checks; generation 1 also takes T % 128 > 1)


def attn_bwd_tc(qkv16, probs16, o32, dO16, ld_do, r, delta_ws, wpart, dqkv16, *, P, B, T, H, splits=1, positive_only=True,
                need_dqkv=True, generation=None):
    """tcgen05 version of attn_bwd (T <= 272)."""
    assert qkv16.dtype == torch.float16 and probs16.dtype == torch.float16
    # algorithmic minimum: G = dO V^T always; dQ = dS K, dK = dS^T Q, dV = A^T dO when the block below needs them
    flops = (8.0 if need_dqkv else 2.0) * T * T * 64 * H * P * B
    gen = ATTN_BWD_GENERATION if generation is None else generation
    if gen in (2, 3) and (T <= 128 or T % 128 <= 1):
        name = f"semabs_attn_bwd_tc{gen}"
        CALL_PROFILE.note(name, flops=flops)
        check(
            getattr(lib(), name)(
                ptr(qkv16), i32(qkv16.stride(0)), ptr(probs16), i32(probs16.shape[-1]), ptr(o32), ptr(dO16), i32(ld_do), ptr(r),
                ptr(delta_ws), ptr(wpart), ptr(dqkv16), i32(P), i32(B), i32(T), i32(H), i32(splits), i32(int(positive_only)),
                i32(int(need_dqkv)), stream_ptr(),
            )
        )
        return
    CALL_PROFILE.note("semabs_attn_bwd_tc", flops=flops)
    check(
        lib().semabs_attn_bwd_tc(
            ptr(qkv16), i32(qkv16.stride(0)), ptr(probs16), i32(probs16.shape[-1]), ptr(o32), ptr(dO16), i32(ld_do), ptr(r),
            ptr(delta_ws), ptr(wpart), ptr(dqkv16), i32(P), i32(B), i32(T), i32(H), i32(splits), i32(int(positive_only)),
            i32(int(need_dqkv)), stream_ptr(),
        )
    )


def attn_bwd_cls(qkv16, probs16, dO16_cls, ld_do, r, wpart, dqkv16, *, P, B, T, H, splits=1, positive_only=True,
                 need_dqkv=True):
    check(
        lib().semabs_attn_bwd_cls(
            ptr(qkv16), i32(qkv16.stride(0)), ptr(probs16), i32(probs16.shape[-1]), ptr(dO16_cls), i32(ld_do), ptr(r), ptr(wpart), ptr(dqkv16),
            i32(P), i32(B), i32(T), i32(H), i32(splits), i32(int(positive_only)), i32(int(need_dqkv)), stream_ptr(),
        )
    )


def clip_logit_seed(f, W, *, B, P, E, logits=None, seed16=None, splits=1):
    check(lib().semabs_clip_logit_seed(ptr(f), ptr(W), ptr(logits), ptr(seed16), i32(B), i32(P), i32(E), i32(splits), stream_ptr()))


def rollout_init(r, PB, T):
    check(lib().semabs_rollout_init(ptr(r), i32(PB), i32(T), stream_ptr()))


def rollout_update(r, wpart, PB, H, T):
    check(lib().semabs_rollout_update(ptr(r), ptr(wpart), i32(PB), i32(H), i32(T), stream_ptr()))


def text_embed(tokens, table, pos, x, n_texts, ctx, d):
    assert tokens.dtype == torch.int32
    check(lib().semabs_text_embed(ptr(tokens), ptr(table), ptr(pos), ptr(x), i32(n_texts), i32(ctx), i32(d), stream_ptr()))


def zeroshot_weights(feat, W, n_classes, n_templates, E):
    check(lib().semabs_zeroshot_weights(ptr(feat), ptr(W), i32(n_classes), i32(n_templates), i32(E), stream_ptr()))


def tile_assemble(rel, tile_desc, size_order, H, W, out):
    P, n, g, _ = rel.shape
    assert rel.is_contiguous() and tile_desc.dtype == torch.int32 and size_order.dtype == torch.int32
    check(lib().semabs_tile_assemble(ptr(rel), ptr(tile_desc), i32(n), ptr(size_order), i32(size_order.numel()), i32(g),
                                     i32(H), i32(W), i32(P), ptr(out), stream_ptr()))
    return out


def tile_preprocess(images_u8, tiles, coef, bounds, out, mean, std):
    """images_u8 [n,H,W,3] uint8, tiles int32 [n_tiles,5], coef int32 [n_sizes,R,kmax], bounds int32 [n_sizes,R,2] (all on
    the device) -> out [n_tiles,3,R,R] fp32; mean / std: 3 python floats each."""
    n_img, H, W, _ = images_u8.shape
    assert images_u8.dtype == torch.uint8 and images_u8.is_contiguous() and tiles.dtype == torch.int32 and tiles.is_contiguous()
    assert coef.dtype == torch.int32 and bounds.dtype == torch.int32 and out.dtype == torch.float32 and out.is_contiguous()
    n_sizes, R, kmax = coef.shape
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    check(lib().semabs_tile_preprocess(ptr(images_u8), i32(n_img), i32(H), i32(W), ptr(tiles), i32(tiles.shape[0]), ptr(coef),
                                       ptr(bounds), i32(n_sizes), i32(kmax), i32(R), m3, s3, ptr(out), stream_ptr()))
    return out


def color_jitter(img_u8: torch.Tensor, fn_idx, brightness, contrast, saturation, hue) -> torch.Tensor:
    """torchvision.transforms.ColorJitter.forward on a uint8 HWC device image with the parameters `ColorJitter.get_params`
    drew (CLIP/clip/__init__.py:55-57,246-247): the four operations in the drawn order, each one launch of
    semabs_color_jitter_op.  Returns a new [H,W,3] uint8 tensor."""
    assert img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.shape[2] == 3
    cur = img_u8.contiguous()
    npix = cur.shape[0] * cur.shape[1]
    scratch = torch.empty(1, dtype=torch.int64, device=cur.device)
    factors = {0: brightness, 1: contrast, 2: saturation, 3: hue}
    for fn_id in [int(i) for i in fn_idx]:
        f = factors[fn_id]
        if f is None:
            continue
        nxt = torch.empty_like(cur)
        check(lib().semabs_color_jitter_op(ptr(cur), ptr(nxt), _i64(npix), i32(fn_id), f32(f), ptr(scratch), stream_ptr()))
        cur = nxt
    return cur if cur is not img_u8 else img_u8.clone()


def flip_average(rel, rel_flipped):
    g = rel.shape[-1]
    assert rel.is_contiguous() and rel_flipped.is_contiguous() and rel.shape == rel_flipped.shape
    check(lib().semabs_flip_average(ptr(rel), ptr(rel_flipped), _i64(rel.numel() // (g * g)), i32(g), stream_ptr()))
    return rel


# ---------------------------------------------------------------------------------------------------------
# 3-D UNet stages
# ---------------------------------------------------------------------------------------------------------
CONV_3X3X3, CONV_1X1X1, CONV_TRANSPOSE_PARITY = 0, 1, 2


def conv3d(x16, w16, *, kind, N, D, H, W, C_in, C_out, a_splits=1, w_splits=1, parity=0, precise=False, bias=None,
           residual=None, relu=False, out32=None, out16=None, o16_splits=1, stats=None, groups=0):
    if CALL_PROFILE.on:
        # taps per output voxel: 27 / 1 / the parity class's tap list (1,2,2,2,4,4,4,8: 27 over the 8 classes);
        # kinds 2 / 3 produce D*H*W outputs per launch (one parity class)
        taps = {CONV_3X3X3: 27, CONV_1X1X1: 1}.get(kind, (1 << bin(parity).count("1")))
        CALL_PROFILE.note("semabs_conv3d", flops=2.0 * N * D * H * W * taps * C_in * C_out)
    check(
        lib().semabs_conv3d(
            ptr(x16), i32(a_splits), ptr(w16), i32(w_splits), i32(kind), i32(parity), i32(N), i32(D), i32(H), i32(W),
            i32(C_in), i32(C_out), i32(int(precise)), ptr(bias), ptr(residual), i32(int(relu)), ptr(out32), ptr(out16),
            i32(o16_splits), ptr(stats), i32(groups), stream_ptr(),
        )
    )


def conv_transpose3d_s2(x16, w16, *, N, D, H, W, C_in, C_out, a_splits, w_splits, precise, bias=None, residual=None, out32=None,
                        out_planar=None, stats=None, groups=0):
    """Transposed convolution (k3 s2 p1, output 2x) with all eight parity classes in one launch; C_out % 32 == 0, C_in % 64 == 0."""
    CALL_PROFILE.note("semabs_conv_transpose3d_s2", flops=2.0 * N * D * H * W * 27 * C_in * C_out)
    check(
        lib().semabs_conv_transpose3d_s2(
            ptr(x16), i32(a_splits), ptr(w16), i32(w_splits), i32(N), i32(D), i32(H), i32(W), i32(C_in), i32(C_out), i32(int(precise)),
            ptr(bias), ptr(residual), ptr(out32), ptr(out_planar), ptr(stats), i32(groups), stream_ptr(),
        )
    )


def ncdhw_to_ndhwc(x, y, *, N, S, C, Cpad, groups=1, stats=None):
    CALL_PROFILE.note("semabs_ncdhw_to_ndhwc", bytes=N * S * (C + Cpad) * 4)
    check(lib().semabs_ncdhw_to_ndhwc(ptr(x), ptr(y), i32(N), _i64(S), i32(C), i32(Cpad), i32(groups), ptr(stats), stream_ptr()))


def ncdhw_to_planar(x, y16, *, N, S, C, Cpad, groups=1, stats=None):
    CALL_PROFILE.note("semabs_ncdhw_to_planar", bytes=N * S * (C * 4 + Cpad * 4))
    check(lib().semabs_ncdhw_to_planar(ptr(x), ptr(y16), i32(N), _i64(S), i32(C), i32(Cpad), i32(groups), ptr(stats), stream_ptr()))


def ndhwc_to_ncdhw(x, y, *, N, S, C):
    CALL_PROFILE.note("semabs_ndhwc_to_ncdhw", bytes=N * S * C * 8)
    check(lib().semabs_ndhwc_to_ncdhw(ptr(x), ptr(y), i32(N), _i64(S), i32(C), stream_ptr()))


def final_conv1x1_ncdhw(x16, w, bias, y, *, N, S, C_in, C_out, splits):
    """final 1x1x1 convolution + channels-last -> NCDHW in one pass (see semabs_final_conv1x1_ncdhw)."""
    CALL_PROFILE.note("semabs_final_conv1x1_ncdhw", bytes=N * S * (2 * splits * C_in + 4 * C_out), flops=2.0 * N * S * C_in * C_out)
    check(lib().semabs_final_conv1x1_ncdhw(ptr(x16), i32(splits), ptr(w), ptr(bias), ptr(y), i32(N), _i64(S), i32(C_in), i32(C_out),
                                           stream_ptr()))


def groupnorm_apply(x, stats, gamma, beta, y16, *, N, S, C, C_real, groups, splits=1, planar=False):
    CALL_PROFILE.note("semabs_groupnorm_apply", bytes=N * S * C * (4 + 2 * splits))
    check(lib().semabs_groupnorm_apply(ptr(x), ptr(stats), ptr(gamma), ptr(beta), ptr(y16), i32(N), _i64(S), i32(C),
                                       i32(C_real), i32(groups), i32(splits), i32(int(planar)), stream_ptr()))


def conv3d_halo(x16_planar, w_img, *, N, D, H, W, C_in, C_out, a_splits=1, w_splits=1, precise=False, residual=None,
                relu=False, out32=None, out16=None, o16_splits=1, stats=None, groups=0):
    CALL_PROFILE.note("semabs_conv3d_halo", flops=2.0 * N * D * H * W * 27 * C_in * C_out)
    check(
        lib().semabs_conv3d_halo(
            ptr(x16_planar), i32(a_splits), ptr(w_img), i32(w_splits), i32(N), i32(D), i32(H), i32(W), i32(C_in),
            i32(C_out), i32(int(precise)), ptr(residual), i32(int(relu)), ptr(out32), ptr(out16), i32(o16_splits),
            ptr(stats), i32(groups), stream_ptr(),
        )
    )


def conv3d_halo_fused(x16_planar, w_img, *, D, H, C_in, a_splits, w_splits, precise, bias_cls=None, residual=None, res_planar=None,
                      relu=False, out32=None, out16=None, o16_splits=1, out_planar=None, stats=None, groups=0):
    """One sample, C_out = 32, CTA-pair kernel with the consumed tensor's GroupNorm folded in (see semabs_conv3d_halo_fused)."""
    CALL_PROFILE.note("semabs_conv3d_halo_fused", flops=2.0 * D * H * 128 * 27 * C_in * 32)
    check(
        lib().semabs_conv3d_halo_fused(
            ptr(x16_planar), i32(a_splits), ptr(w_img), i32(w_splits), i32(D), i32(H), i32(C_in), i32(int(precise)), ptr(bias_cls),
            ptr(residual), ptr(res_planar), i32(int(relu)), ptr(out32), ptr(out16), i32(o16_splits), ptr(out_planar), ptr(stats),
            i32(groups), stream_ptr(),
        )
    )


def fold_groupnorm_halo(w, gamma, beta, stats, w_img, bias_cls, *, N, S, groups):
    """Per-sample halo weight images of W * gamma * rstd + border-class bias tables, one launch (see semabs_fold_groupnorm_halo)."""
    assert stats.dim() == 3 and stats.is_contiguous() and stats.shape[0] == N
    check(lib().semabs_fold_groupnorm_halo(ptr(w), ptr(gamma), ptr(beta), ptr(stats), i32(stats.shape[1] * stats.shape[2]), i32(N),
                                           _i64(S), i32(groups), ptr(w_img), ptr(bias_cls), stream_ptr()))


def set_halo_pair(enable: bool) -> None:
    """CTA-pair variant of the halo-resident convolution on / off (default on); see semabs_set_halo_pair."""
    check(lib().semabs_set_halo_pair(i32(int(enable))))


def pack_halo_weights(w: torch.Tensor, splits: int) -> torch.Tensor:
    """conv.weight [Co, Ci, 3,3,3] fp32 -> resident UMMA no-swizzle core-matrix images for semabs_conv3d_halo, fp16:
    [Co/16 halves][27 taps][Ci/16 k-blocks][N/8 groups][2 k-chunks][8 rows][8 elems], where the N rows of a half are
    its 16 output channels of W_hi followed (splits == 2) by the same 16 channels of W_lo = fp16(W - W_hi)."""
    co, ci = w.shape[:2]
    w = w.detach().float().reshape(co, ci, 27)
    hi = w.half()
    halves = []
    for h in range(co // 16):
        rows = [hi[h * 16 : (h + 1) * 16]]
        if splits == 2:
            rows.append((w - hi.float()).half()[h * 16 : (h + 1) * 16])
        b = torch.cat(rows, dim=0)  # [N, ci, 27]
        n = b.shape[0]
        t = b.permute(2, 0, 1).reshape(27, n // 8, 8, ci // 16, 2, 8)  # tap, g, r, kb, kc, e
        halves.append(t.permute(0, 3, 1, 4, 2, 5).contiguous())  # tap, kb, g, kc, r, e
    return torch.stack(halves, dim=0).contiguous()


def maxpool3d_2(x, y, *, N, D, H, W, C, groups=1, stats=None):
    CALL_PROFILE.note("semabs_maxpool3d_2", bytes=N * D * H * W * C * 4 * 1.125)
    check(lib().semabs_maxpool3d_2(ptr(x), ptr(y), i32(N), i32(D), i32(H), i32(W), i32(C), i32(groups), ptr(stats), stream_ptr()))


# ---------------------------------------------------------------------------------------------------------
# point <-> voxel stages
# ---------------------------------------------------------------------------------------------------------
_cf, _ci = C.c_float, C.c_int32


def _host3(vals, ctype):
    return (ctype * 3)(*vals)


def points_to_voxels(xyz, feat, grid, *, N, npts, F, xyz_div, mlp, C_out, vol, cnt, Cpad, groups=1, stats=None):
    """grid = (neg_lc[3], scale[3], shape[3]) host values; mlp = None or (w1t,b1,w2t,b2,w3t,b3, hidden)."""
    neg_lc, scale, shape = grid
    if stats is not None and groups > 1 and (C_out % groups != 0 or (C_out // groups) % 2 != 0):
        # the finalize kernel accumulates the statistics channel pair by channel pair: a pair must not straddle two groups
        raise ValueError(f"points_to_voxels: GroupNorm statistics need an even number of channels per group (C={C_out}, groups={groups})")
    w = mlp[:6] if mlp is not None else (None,) * 6
    hidden = mlp[6] if mlp is not None else 0
    check(
        lib().semabs_points_to_voxels(
            ptr(xyz), i32(xyz_div), ptr(feat), i32(N), i32(npts), i32(F), i32(int(mlp is not None)), i32(hidden), i32(C_out),
            ptr(w[0]), ptr(w[1]), ptr(w[2]), ptr(w[3]), ptr(w[4]), ptr(w[5]), _host3(neg_lc, _cf),
            _host3(scale, _cf), _host3(shape, _ci), ptr(vol), ptr(cnt), i32(Cpad), i32(groups), ptr(stats), stream_ptr(),
        )
    )


def sample_decode(vol0, vol1, C0, query, grid, *, N, nq, concat_xyz, w1t, b1, w2t, b2, Hs, out_dim, out, emb=None,
                  temperature=1.0):
    neg_lc, scale, shape = grid
    check(
        lib().semabs_sample_decode(
            ptr(vol0), ptr(vol1), i32(C0), ptr(query), i32(N), i32(nq), _host3(neg_lc, _cf), _host3(scale, _cf),
            _host3(shape, _ci), i32(int(concat_xyz)), ptr(w1t), ptr(b1), ptr(w2t), ptr(b2), i32(Hs), i32(out_dim),
            ptr(emb), f32(temperature), ptr(out), stream_ptr(),
        )
    )


# ---------------------------------------------------------------------------------------------------------
# UNet backward stages (unet_bwd.cu)
# ---------------------------------------------------------------------------------------------------------
CONV_TRANSPOSE_ADJOINT = 3
WGRAD_KB = 64  # voxels per pipeline stage of the weight-gradient kernel: padded volumes are sized in multiples of it


def absmax_f32(x, amax_slot):
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() % 4 == 0
    check(lib().semabs_absmax_f32(ptr(x), _i64(x.numel()), ptr(amax_slot), stream_ptr()))


def unet_bwd_pack(g, *, N, D, H, W, C, g_scale=None, amax=None, mask=None, pad16=None, Cp=0, parity=False, op16=None,
                  op_layout=1, op_splits=1, scale_out=None):
    check(lib().semabs_unet_bwd_pack(ptr(g), ptr(g_scale), ptr(amax), ptr(mask), i32(N), i32(D), i32(H), i32(W), i32(C),
                                     ptr(pad16), i32(Cp), i32(int(parity)), ptr(op16), i32(op_layout), i32(op_splits),
                                     ptr(scale_out), stream_ptr()))


def groupnorm_apply_padded(x, stats, gamma, beta, pad16, *, N, D, H, W, C, C_real, groups, Cp):
    check(lib().semabs_groupnorm_apply_padded(ptr(x), ptr(stats), ptr(gamma), ptr(beta), ptr(pad16), i32(N), i32(D), i32(H),
                                              i32(W), i32(C), i32(C_real), i32(groups), i32(Cp), stream_ptr()))


def conv3d_wgrad(A16, B16, *, lda, Ca, Ca_real, ldb, Cb, Cb_real, nvox, seg_off, seg_ntaps, seg_sh, seg_slot, nslots,
                 slot_k, KT, workspace, scale, grad, accumulate=False):
    """seg_off / seg_ntaps: python lists (one entry per segment); seg_sh / seg_slot: lists of 3-lists; slot_k: device int32."""
    nseg = len(seg_off)
    off = (C.c_int64 * nseg)(*[int(v) for v in seg_off])
    nt = (C.c_int32 * nseg)(*[int(v) for v in seg_ntaps])
    sh = (C.c_int32 * (3 * nseg))(*[int(v) for row in seg_sh for v in row])
    sl = (C.c_int32 * (3 * nseg))(*[int(v) for row in seg_slot for v in row])
    assert slot_k.dtype == torch.int32 and slot_k.numel() >= nslots and grad.dtype == torch.float32 and grad.is_contiguous()
    check(lib().semabs_conv3d_wgrad(ptr(A16), i32(lda), i32(Ca), i32(Ca_real), ptr(B16), i32(ldb), i32(Cb), i32(Cb_real),
                                    _i64(nvox), i32(nseg), off, nt, sh, sl, i32(nslots), ptr(slot_k), i32(KT), ptr(workspace),
                                    _i64(workspace.numel() * workspace.element_size()), ptr(scale), ptr(grad),
                                    i32(int(accumulate)), stream_ptr()))


def groupnorm_bwd_reduce(dy, x, sums, *, N, S, C):
    check(lib().semabs_groupnorm_bwd_reduce(ptr(dy), ptr(x), i32(N), _i64(S), i32(C), ptr(sums), stream_ptr()))


def groupnorm_bwd_apply(dy, x, stats, gamma, sums, dx, *, N, S, C, C_real, groups, dy_scale=None, add=None, add_scale=None,
                        add_mask=None, accumulate=False, amax=None):
    check(lib().semabs_groupnorm_bwd_apply(ptr(dy), ptr(dy_scale), ptr(x), ptr(stats), ptr(gamma), ptr(sums), i32(N), _i64(S),
                                           i32(C), i32(C_real), i32(groups), ptr(add), ptr(add_scale), ptr(add_mask), ptr(dx),
                                           i32(int(accumulate)), ptr(amax), stream_ptr()))


def groupnorm_param_grads(sums, *, N, S, C, C_real, groups=1, stats=None, scale=None, dgamma=None, dbeta=None, accumulate=False):
    check(lib().semabs_groupnorm_param_grads(ptr(sums), ptr(stats), ptr(scale), i32(N), _i64(S), i32(C), i32(C_real),
                                             i32(groups), ptr(dgamma), ptr(dbeta), i32(int(accumulate)), stream_ptr()))


def maxpool3d_2_bwd(g, x, dx, *, N, D, H, W, C, g_scale=None, accumulate=False, amax=None):
    check(lib().semabs_maxpool3d_2_bwd(ptr(g), ptr(g_scale), ptr(x), i32(N), i32(D), i32(H), i32(W), i32(C), ptr(dx),
                                       i32(int(accumulate)), ptr(amax), stream_ptr()))


# ---------------------------------------------------------------------------------------------------------
# point <-> voxel backward (points_bwd.cu)
# ---------------------------------------------------------------------------------------------------------
def outer_reduce_f32(A, B, out, *, R, Cc, P, scale=1.0):
    """out[r, c] += scale * sum_p A[p, r] * B[p, c]; A / B are 2-D fp32 views with unit column stride."""
    assert A.dtype == torch.float32 and B.dtype == torch.float32 and A.stride(1) == 1 and B.stride(1) == 1
    assert out.dtype == torch.float32 and out.stride(-1) == 1
    check(lib().semabs_outer_reduce_f32(ptr(A), i32(A.stride(0)), i32(R), ptr(B), i32(B.stride(0)), i32(Cc), _i64(P),
                                        f32(scale), ptr(out), i32(out.stride(0)), stream_ptr()))


def sample_decode_bwd(vol0, vol1, C0, query, grid, *, N, nq, concat_xyz, w1t, w1, b1, w2t, w2, b2, Hs, out_dim, dout, dvol0,
                      dvol1, scratch, off_h, off_do, off_dp, emb=None, temperature=1.0, demb=None):
    neg_lc, scale, shape = grid
    check(
        lib().semabs_sample_decode_bwd(
            ptr(vol0), ptr(vol1), i32(C0), ptr(query), i32(N), i32(nq), _host3(neg_lc, _cf), _host3(scale, _cf),
            _host3(shape, _ci), i32(int(concat_xyz)), ptr(w1t), ptr(w1), ptr(b1), ptr(w2t), ptr(w2), ptr(b2), i32(Hs),
            i32(out_dim), ptr(emb), f32(temperature), ptr(dout), ptr(dvol0), ptr(dvol1), ptr(demb), ptr(scratch),
            i32(scratch.stride(0)), i32(off_h), i32(off_do), i32(off_dp), stream_ptr(),
        )
    )


def points_to_voxels_bwd(xyz, feat, grid, *, N, npts, F, xyz_div, hidden, C, w1t, b1, w2t, w2, b2, w3, dvol, cnt, Cpad,
                         scratch, off_d3, off_d2, off_d1):
    neg_lc, scale, shape = grid
    check(
        lib().semabs_points_to_voxels_bwd(
            ptr(xyz), i32(xyz_div), ptr(feat), i32(N), i32(npts), i32(F), i32(hidden), i32(C), ptr(w1t), ptr(b1), ptr(w2t),
            ptr(w2), ptr(b2), ptr(w3), _host3(neg_lc, _cf), _host3(scale, _cf), _host3(shape, _ci), ptr(dvol), ptr(cnt),
            i32(Cpad), ptr(scratch), i32(scratch.stride(0)), i32(off_d3), i32(off_d2), i32(off_d1), stream_ptr(),
        )
    )
