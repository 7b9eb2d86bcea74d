"""Thin typed wrappers: torch device tensors in, C-ABI calls out. One function per entry point."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_QUICKGELU, ACT_QUICKGELU_GRAD, GemmEpilogue, check, f32, i32, lib, ptr, stream_ptr


def gemm_f16(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    a_splits: int = 1,
    bias: torch.Tensor | None = None,
    residual: torch.Tensor | None = None,
    aux: torch.Tensor | None = None,
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
    ep.aux = aux.data_ptr() if aux is not None else None
    ep.aux_rows = aux.shape[0] if aux is not None else 0
    ep.ld_aux = aux.stride(0) if aux is not None else 0
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
    check(
        lib().semabs_gemm_f16(
            ptr(a), i32(a.stride(0)), ptr(b), i32(b.stride(0)), i32(M), i32(N), i32(K), i32(a_splits),
            C.byref(ep), stream_ptr(),
        )
    )
    return out_f32 if out_f32 is not None else out_f16


def split_f16(x: torch.Tensor) -> torch.Tensor:
    """[M,K] fp32 -> [M,2K] fp16 (hi | lo). Test helper; product kernels emit the split in their epilogues."""
    hi = x.half()
    lo = (x - hi.float()).half()
    return torch.cat([hi, lo], dim=1).contiguous()
