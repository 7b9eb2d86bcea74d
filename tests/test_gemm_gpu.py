"""tcgen05/TMA GEMM (semabs_gemm_f16) vs torch fp32/fp64 matmul on the same fp16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a16, b16):
    return (a16.double() @ b16.double().t()).float()


@pytest.mark.parametrize(
    "M,N,K",
    [(128, 128, 64), (256, 128, 128), (300, 384, 192), (8224, 3072, 1024), (1000, 64, 640), (77, 32, 512), (4096, 512, 1024)],
)
def test_gemm_plain(M, N, K):
    from semabs_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    b = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).half()
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm_f16(a, b, out_f32=out)
    torch.cuda.synchronize()
    ref = _ref(a, b)
    err = (out - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


def test_gemm_split_is_fp32_accurate():
    from semabs_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(7)
    M, N, K = 520, 256, 1024
    a = torch.randn(M, K, device="cuda", generator=g)
    b = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).half()
    out = torch.empty(M, N, device="cuda")
    ops.gemm_f16(ops.split_f16(a), b, a_splits=2, out_f32=out)
    ref = (a.double() @ b.double().t()).float()
    assert (out - ref).abs().max().item() < 2e-5 * ref.abs().max().item()
    out1 = torch.empty(M, N, device="cuda")
    ops.gemm_f16(a.half(), b, out_f32=out1)
    assert (out1 - ref).abs().max().item() > (out - ref).abs().max().item()


def test_gemm_epilogues():
    from semabs_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 514, 256, 128
    a = torch.randn(M, K, device="cuda", generator=g).half()
    b = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    base = _ref(a, b)

    # bias + column scaling + residual, fp32 and fp16 (split) outputs
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, 2 * N, device="cuda", dtype=torch.float16)
    ops.gemm_f16(a, b, bias=bias, residual=res, out_f32=o32, out_f16=o16, out_f16_splits=2, scale_cols=128, scale=0.125)
    ref = base + bias
    ref[:, :128] *= 0.125
    ref = ref + res
    assert torch.allclose(o32, ref, atol=1e-4, rtol=1e-5)
    recon = o16[:, :N].float() + o16[:, N:].float()
    assert torch.allclose(recon, o32, atol=1e-5, rtol=1e-5)

    # quickgelu: fp32 = pre-activation, fp16 = activation, aux16 = derivative
    u = torch.empty(M, N, device="cuda")
    gq = torch.empty(M, N, device="cuda", dtype=torch.float16)
    gg = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ops.gemm_f16(a, b, bias=bias, act=ops.ACT_QUICKGELU, out_f32=u, out_f16=gq, out_aux16=gg)
    pre = base + bias
    sg = torch.sigmoid(1.702 * pre)
    assert torch.allclose(u, pre, atol=1e-4, rtol=1e-5)
    assert torch.allclose(gq.float(), pre * sg, atol=2e-3, rtol=2e-3)
    assert torch.allclose(gg.float(), sg + 1.702 * pre * sg * (1 - sg), atol=2e-3, rtol=2e-3)

    # dgrad through quickgelu: multiply by a row-broadcast fp16 aux (aux_rows divides M)
    aux = torch.randn(M // 2, N, device="cuda", generator=g).half()
    d = torch.empty(M, N, device="cuda")
    ops.gemm_f16(a, b, aux16=aux, act=ops.ACT_MUL_AUX16, out_f32=d)
    assert torch.allclose(d, base * aux.float().repeat(2, 1), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("S,P,N", [(257, 16, 256), (4 * 257, 5, 512), (1024, 3, 128), (200, 7, 256)])
def test_gemm_aux_rasterised_tile_order(S, P, N):
    """MUL_AUX16 with aux_rows < M visits the P repeats of an aux row block back to back (gemm.cu tile_coords); the
    repeat stride is not a multiple of the 128-row tile, so row blocks straddle repeats — every row must still be
    written exactly once."""
    from semabs_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(S + P)
    M, K = S * P, 192
    a = torch.randn(M, K, device="cuda", generator=g).half()
    b = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).half()
    aux = torch.randn(S, N, device="cuda", generator=g).half()
    d32 = torch.full((M, N), float("nan"), device="cuda")
    d16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
    ops.gemm_f16(a, b, aux16=aux, act=ops.ACT_MUL_AUX16, out_f32=d32, out_f16=d16)
    ref = _ref(a, b) * aux.float().repeat(P, 1)
    assert torch.allclose(d32, ref, atol=1e-4, rtol=1e-4)
    assert torch.allclose(d16.float(), ref, atol=2e-3, rtol=2e-3)


def test_gemm_unaligned_rows_use_narrow_accesses():
    """Output / residual rows that are only 16-byte aligned (a column slice of a wider buffer) take the 128-bit path."""
    from semabs_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 300, 64, 128
    a = torch.randn(M, K, device="cuda", generator=g).half()
    b = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).half()
    wide32 = torch.zeros(M, N + 4, device="cuda")
    res = torch.randn(M, N + 4, device="cuda", generator=g)
    out = wide32[:, 4:]
    assert out.data_ptr() % 32 != 0
    ops.gemm_f16(a, b, residual=res[:, 4:], out_f32=out)
    assert torch.allclose(out, _ref(a, b) + res[:, 4:], atol=1e-4, rtol=1e-4)
    assert wide32[:, :4].abs().max().item() == 0.0


@pytest.mark.parametrize("M,N,K", [(9700, 1024, 512), (38000, 256, 1024), (8224, 3072, 1024)])
def test_gemm_cta_pair_kernel_matches_single_cta_and_reference(M, N, K):
    """Shapes with >= 2 x 148 tiles of 128 x 256 take the CTA-pair kernel (gemm2.cu, tcgen05.mma.cta_group::2, 256 x 256 tile
    per pair; M deliberately not a multiple of 256): every epilogue flavour against fp64 torch and against the single-CTA
    kernel on the same inputs."""
    from semabs_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    b = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    base = _ref(a, b)
    outs = {}
    for pair in (True, False):
        ops.set_gemm_pair(pair)
        try:
            o32 = torch.full((M, N), float("nan"), device="cuda")
            o16 = torch.full((M, 2 * N), float("nan"), device="cuda", dtype=torch.float16)
            ops.gemm_f16(a, b, bias=bias, residual=res, out_f32=o32, out_f16=o16, out_f16_splits=2, scale_cols=128, scale=0.125)
            gq = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
            gg = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
            ops.gemm_f16(a, b, bias=bias, act=ops.ACT_QUICKGELU, out_f16=gq, out_aux16=gg)
            a2 = ops.split_f16(a.float() * 1.0003)
            osplit = torch.full((M, N), float("nan"), device="cuda")
            ops.gemm_f16(a2, b, a_splits=2, out_f32=osplit)
            torch.cuda.synchronize()
        finally:
            ops.set_gemm_pair(True)
        outs[pair] = (o32, o16, gq, gg, osplit)
    o32, o16, gq, gg, osplit = outs[True]
    ref = base + bias
    ref[:, :128] *= 0.125
    ref = ref + res
    assert torch.allclose(o32, ref, atol=2e-4, rtol=1e-5)
    assert torch.allclose(o16[:, :N].float() + o16[:, N:].float(), o32, atol=1e-5, rtol=1e-5)
    pre = base + bias
    sg = torch.sigmoid(1.702 * pre)
    assert torch.allclose(gq.float(), pre * sg, atol=3e-3, rtol=2e-3)
    assert torch.allclose(gg.float(), sg + 1.702 * pre * sg * (1 - sg), atol=3e-3, rtol=2e-3)
    ref_split = ((a.float() * 1.0003).double() @ b.double().t()).float()
    assert (osplit - ref_split).abs().max().item() < 3e-5 * ref_split.abs().max().item()
    for x, y in zip(outs[True], outs[False]):  # same K order of accumulation in both kernels: identical bits expected
        assert torch.equal(x, y) or torch.allclose(x.float(), y.float(), atol=1e-5, rtol=1e-5)


def test_gemm_cta_pair_aux_raster():
    from semabs_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(99)
    S, P, N, K = 5 * 257, 8, 1024, 256
    M = S * P
    a = torch.randn(M, K, device="cuda", generator=g).half()
    b = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).half()
    aux = torch.randn(S, N, device="cuda", generator=g).half()
    d16 = torch.full((M, 2 * N), float("nan"), device="cuda", dtype=torch.float16)
    ops.gemm_f16(a, b, aux16=aux, act=ops.ACT_MUL_AUX16, out_f16=d16, out_f16_splits=2)
    ref = _ref(a, b) * aux.float().repeat(P, 1)
    got = d16[:, :N].float() + d16[:, N:].float()
    assert not torch.isnan(got).any()
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)
