"""Unit parity of the non-GEMM ViT kernels against plain torch fp32 on the same device."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


def _recon(y16, d, splits):
    return y16[:, :d].float() + (y16[:, d:].float() if splits == 2 else 0)


@pytest.mark.parametrize("splits", [1, 2])
def test_layernorm_fwd_bwd(splits):
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(0)
    M, d, P = 37, 768, 3
    x = torch.randn(M, d, device=dev, generator=g) * 2 + 0.5
    gamma = 1 + 0.1 * torch.randn(d, device=dev, generator=g)
    beta = 0.1 * torch.randn(d, device=dev, generator=g)
    y32 = torch.empty(M, d, device=dev)
    y16 = torch.empty(M, splits * d, device=dev, dtype=torch.float16)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    ops.layernorm_fwd(x, gamma, beta, M=M, d=d, y32=y32, y16=y16, mean=mean, rstd=rstd, splits=splits)
    ref = F.layer_norm(x, (d,), gamma, beta, 1e-5)
    assert torch.allclose(y32, ref, atol=2e-6, rtol=1e-5)
    tol = 1e-6 if splits == 2 else 2e-3
    assert torch.allclose(_recon(y16, d, splits), ref, atol=tol * 4, rtol=tol)

    dy = torch.randn(P * M, d, device=dev, generator=g)
    dres = torch.randn(P * M, d, device=dev, generator=g)
    dx = torch.empty(P * M, d, device=dev)
    dx16 = torch.empty(P * M, splits * d, device=dev, dtype=torch.float16)
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, M=P * M, d=d, x_rows=M, dres=dres, dx16=dx16, splits=splits)
    xr = x.repeat(P, 1).requires_grad_(True)
    F.layer_norm(xr, (d,), gamma, beta, 1e-5).backward(dy)
    refdx = xr.grad + dres
    assert torch.allclose(dx, refdx, atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("T,H,causal", [(50, 12, False), (257, 16, False), (77, 8, True)])
def test_attn_fwd(T, H, causal):
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(T)
    B, d = 3, H * 64
    qkv = torch.randn(B * T, 3 * d, device=dev, generator=g)
    qkv[:, :d] *= 0.125 * 2.0
    probs = torch.empty(B * H, T, T, device=dev)
    o32 = torch.empty(B * T, d, device=dev)
    o16 = torch.empty(B * T, 2 * d, device=dev, dtype=torch.float16)
    ops.attn_fwd(qkv, B=B, T=T, H=H, probs=probs, o32=o32, o16=o16, causal=causal, splits=2)
    q, k, v = (t.reshape(B, T, H, 64).permute(0, 2, 1, 3) for t in qkv.view(B, T, 3 * d).chunk(3, -1))
    s = q @ k.transpose(-1, -2)
    if causal:
        s = s + torch.full((T, T), float("-inf"), device=dev).triu_(1)
    a = s.softmax(-1)
    o = (a @ v).permute(0, 2, 1, 3).reshape(B * T, d)
    assert torch.allclose(probs.view(B, H, T, T), a, atol=1e-6, rtol=1e-4)
    assert torch.allclose(o32, o, atol=1e-5, rtol=1e-4)
    assert torch.allclose(_recon(o16, d, 2), o, atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("T,H,impl", [(50, 12, "mma"), (257, 16, "mma"), (50, 12, "tc"), (257, 16, "tc"), (197, 12, "tc"),
                                      (128, 4, "tc"), (130, 4, "tc"), (50, 12, "tc1"), (257, 16, "tc1"), (128, 4, "tc1"),
                                      (129, 2, "tc"), (256, 2, "tc"), (50, 12, "tc2"), (257, 16, "tc2"), (129, 2, "tc2"),
                                      (16, 2, "tc"), (8, 2, "tc")])
def test_attn_bwd(T, H, impl):
    """impl: "mma" = mma.sync kernels; "tc" = the product path (third-generation tcgen05 kernels, vit_attn_bwd3.cu, where
    T % 128 <= 1; first generation otherwise); "tc2" / "tc1" = second / first generation forced (cross-checks)."""
    import functools

    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(T + 1)
    B, P, d = 2, 3, H * 64
    qkv = torch.randn(B * T, 3 * d, device=dev, generator=g)
    qkv[:, :d] *= 0.125 * 1.5
    probs = torch.empty(B * H, T, T, device=dev)
    Tp = (T + 15) // 16 * 16
    probs16 = torch.full((B * H, T, Tp), float("nan"), device=dev, dtype=torch.float16)
    o32 = torch.empty(B * T, d, device=dev)
    ops.attn_fwd(qkv, B=B, T=T, H=H, probs=probs, probs16=probs16, o32=o32)
    assert torch.equal(probs16[..., :T], probs.half()) and (probs16[..., T:] == 0).all()
    qkv16 = qkv.half()
    dO = torch.randn(P * B * T, d, device=dev, generator=g).half()
    r = torch.rand(P * B, T, device=dev, generator=g)
    delta = torch.empty(P * B * H, T, device=dev)
    wpart = torch.full((P * B * H, T), float("nan"), device=dev)
    dqkv16 = torch.full((P * B * T, 2 * 3 * d), float("nan"), device=dev, dtype=torch.float16)
    tc = {"tc": ops.attn_bwd_tc, "tc1": functools.partial(ops.attn_bwd_tc, generation=1), "tc2": functools.partial(ops.attn_bwd_tc, generation=2)}
    bwd = tc.get(impl, ops.attn_bwd)
    bwd(qkv16, probs16, o32, dO, d, r, delta, wpart, dqkv16, P=P, B=B, T=T, H=H, splits=2, positive_only=True)
    torch.cuda.synchronize()
    if impl in tc:  # relevance-only call (last dense block): same wpart, nothing else touched
        w2 = torch.full_like(wpart, float("nan"))
        tc[impl](qkv16, probs16, o32, dO, d, r, delta, w2, None, P=P, B=B, T=T, H=H, splits=2, positive_only=True,
                 need_dqkv=False)
        assert torch.equal(w2, wpart)

    # torch reference (fp32, same fp16-rounded dO)
    dOf = dO.float().view(P, B, T, H, 64).permute(0, 1, 3, 2, 4)  # P,B,H,T,hd
    q, k, v = (t.reshape(B, T, H, 64).permute(0, 2, 1, 3) for t in qkv.view(B, T, 3 * d).chunk(3, -1))
    A = probs.view(B, H, T, T)
    G = dOf @ v.transpose(-1, -2)[None]  # P,B,H,T,T
    cam = (G * A[None]).clamp(min=0)
    w_ref = torch.einsum("pbi,pbhij->pbhj", r.view(P, B, T), cam) / H
    assert torch.allclose(wpart.view(P, B, H, T), w_ref, atol=2e-3 * w_ref.abs().max().item(), rtol=2e-3)
    D = (G * A[None]).sum(-1, keepdim=True)
    dS = A[None] * (G - D)
    dq = (dS @ k[None]) * 0.125
    dk = dS.transpose(-1, -2) @ q[None]
    dv = A[None].transpose(-1, -2).expand(P, -1, -1, -1, -1) @ dOf
    ref = torch.cat([t.permute(0, 1, 3, 2, 4).reshape(P * B * T, d) for t in (dq, dk, dv)], dim=1)
    got = dqkv16[:, : 3 * d].float() + dqkv16[:, 3 * d :].float()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 4e-3 * scale, ((got - ref).abs().max().item(), scale)


@pytest.mark.parametrize("B,T,H,P", [(12, 257, 16, 3), (40, 50, 12, 5), (7, 257, 16, 16), (29, 257, 16, 5), (64, 257, 16, 1),
                                     (29, 257, 16, 2), (3, 257, 16, 33)])
def test_attn_bwd_pipelined_matches_second_generation_over_unit_boundaries(B, T, H, P):
    """More units than SMs, so every CTA of the chunk-pipelined passes (vit_attn_bwd3.cu) crosses unit boundaries (K / V double
    buffer, probability reload, accumulator hand-over).  The first version of these kernels did the same arithmetic in the
    same order as the second generation and matched it bit for bit (commit c8e4e9d); since then dS is formed with one packed
    fp16 multiply (one more rounding of a factor) and the relevance sum runs in two chains, so the comparison carries the
    fp16-operand tolerance; a scheduling error (stale chunk, wrong accumulator, wrong label) would be O(1)."""
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(B * T + P)
    d = H * 64
    qkv = torch.randn(B * T, 3 * d, device=dev, generator=g)
    qkv[:, :d] *= 0.125 * 1.5
    Tp = (T + 15) // 16 * 16
    probs16 = torch.empty(B * H, T, Tp, device=dev, dtype=torch.float16)
    o32 = torch.empty(B * T, d, device=dev)
    ops.attn_fwd(qkv, B=B, T=T, H=H, probs=None, probs16=probs16, o32=o32)
    qkv16 = qkv.half()
    dO = torch.randn(P * B * T, d, device=dev, generator=g).half()
    r = torch.rand(P * B, T, device=dev, generator=g)
    out = {}
    for gen in (2, 3):
        delta = torch.empty(P * B * H, T, device=dev)
        wpart = torch.full((P * B * H, T), float("nan"), device=dev)
        dqkv16 = torch.full((P * B * T, 2 * 3 * d), float("nan"), device=dev, dtype=torch.float16)
        ops.attn_bwd_tc(qkv16, probs16, o32, dO, d, r, delta, wpart, dqkv16, P=P, B=B, T=T, H=H, splits=2, positive_only=True, generation=gen)
        w2 = torch.full_like(wpart, float("nan"))
        ops.attn_bwd_tc(qkv16, probs16, o32, dO, d, r, delta, w2, None, P=P, B=B, T=T, H=H, splits=2, positive_only=True,
                        need_dqkv=False, generation=gen)
        torch.cuda.synchronize()
        assert torch.equal(w2, wpart)
        out[gen] = (wpart, dqkv16)
    assert torch.isfinite(out[3][0]).all() and torch.isfinite(out[3][1].float()).all()
    w2_, w3_ = out[2][0], out[3][0]
    assert (w2_ - w3_).abs().max().item() < 1e-5 * w2_.abs().max().item()
    g2 = out[2][1][:, : 3 * d].float() + out[2][1][:, 3 * d :].float()
    g3 = out[3][1][:, : 3 * d].float() + out[3][1][:, 3 * d :].float()
    err = (g2 - g3).abs().max().item() / g2.abs().max().item()
    print(f"attention backward gen 3 vs gen 2 (B={B}, T={T}, H={H}, P={P}): max|d|/max = {err:.2e}")
    assert err < 1.5e-3


def test_logit_seed():
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(5)
    B, P, E = 5, 3, 512
    f = torch.randn(B, E, device=dev, generator=g)
    W = torch.randn(E, P, device=dev, generator=g) / E**0.5
    logits = torch.empty(B, P, device=dev)
    seed16 = torch.empty(P * B, 2 * E, device=dev, dtype=torch.float16)
    ops.clip_logit_seed(f, W, B=B, P=P, E=E, logits=logits, seed16=seed16, splits=2)
    fr = f.clone().requires_grad_(True)
    lg = 100.0 * (fr / fr.norm(dim=-1, keepdim=True)) @ W
    assert torch.allclose(logits, lg, atol=1e-4, rtol=1e-5)
    for p in range(P):
        (gr,) = torch.autograd.grad(lg[:, p].sum(), fr, retain_graph=True)
        got = _recon(seed16[p * B : (p + 1) * B], E, 2)
        assert torch.allclose(got, gr, atol=1e-5, rtol=1e-3)


@pytest.mark.parametrize("T,H", [(50, 12), (257, 16)])
def test_attn_bwd_class_token_only_matches_dense(T, H):
    """Last-block shortcut: dO non-zero at the class-token row only -> semabs_attn_bwd_cls == semabs_attn_bwd."""
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(T + 7)
    B, P, d = 2, 3, H * 64
    qkv = torch.randn(B * T, 3 * d, device=dev, generator=g)
    qkv[:, :d] *= 0.125 * 1.5
    Tp = (T + 15) // 16 * 16
    probs16 = torch.empty(B * H, T, Tp, device=dev, dtype=torch.float16)
    o32 = torch.empty(B * T, d, device=dev)
    ops.attn_fwd(qkv, B=B, T=T, H=H, probs16=probs16, o32=o32)
    qkv16 = qkv.half()
    dO_cls = torch.randn(P * B, d, device=dev, generator=g).half()
    dO = torch.zeros(P * B, T, d, device=dev, dtype=torch.float16)
    dO[:, 0] = dO_cls
    r = torch.rand(P * B, T, device=dev, generator=g)
    delta = torch.empty(P * B * H, T, device=dev)
    w_ref = torch.empty(P * B * H, T, device=dev)
    dq_ref = torch.empty(P * B * T, 2 * 3 * d, device=dev, dtype=torch.float16)
    ops.attn_bwd(qkv16, probs16, o32, dO.view(-1, d), d, r, delta, w_ref, dq_ref, P=P, B=B, T=T, H=H, splits=2)
    w = torch.full_like(w_ref, float("nan"))
    dq = torch.full_like(dq_ref, float("nan"))
    ops.attn_bwd_cls(qkv16, probs16, dO_cls, d, r, w, dq, P=P, B=B, T=T, H=H, splits=2)
    torch.cuda.synchronize()
    assert torch.allclose(w, w_ref, atol=2e-3 * w_ref.abs().max().item(), rtol=2e-3)
    a = dq[:, : 3 * d].float() + dq[:, 3 * d :].float()
    b = dq_ref[:, : 3 * d].float() + dq_ref[:, 3 * d :].float()
    assert not torch.isnan(a).any()
    assert (a - b).abs().max().item() < 4e-3 * b.abs().max().item()


def test_tcgen05_tmem_a_operand_and_mn_major_b():
    """Known-answer test of the two operand forms vit_attn_tc.cu adds to the GEMM's: A read from TMEM (packed fp16
    pairs stored with tcgen05.st) and an MN-major 128B-swizzled B tile."""
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(5)
    for Kd in (16, 64, 256):
        A = torch.randn(128, Kd, device=dev, generator=g).half()
        Bm = torch.randn(Kd, 64, device=dev, generator=g).half()
        D = torch.full((128, 64), float("nan"), device=dev)
        ops.selftest_ts_mma(A, Bm, D)
        ref = (A.double() @ Bm.double()).float()
        assert torch.allclose(D, ref, atol=1e-3, rtol=1e-4), (Kd, (D - ref).abs().max().item())


@pytest.mark.parametrize("T,H,causal,splits", [(50, 12, False, 2), (257, 16, False, 2), (77, 8, True, 2), (257, 16, False, 1),
                                               (197, 12, False, 2), (128, 4, False, 2), (129, 4, True, 2),
                                               # rows beyond the last full 128-row tile ride on SIMT warps (1, 2 or 3 of them)
                                               (130, 4, False, 1), (131, 4, True, 2), (259, 2, False, 2), (132, 4, False, 2)])
def test_attn_fwd_tc(T, H, causal, splits):
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(T)
    B, d = 3, H * 64
    qkv = torch.randn(B * T, 3 * d, device=dev, generator=g)
    qkv[:, :d] *= 0.125 * 2.0
    if splits == 1:
        qkv = qkv.half().float()
    qkv16 = ops.split_f16(qkv) if splits == 2 else qkv.half()
    ldp = (T + 15) // 16 * 16
    probs16 = torch.full((B * H, T, ldp), float("nan"), device=dev, dtype=torch.float16)
    o32 = torch.full((B * T, d), float("nan"), device=dev)
    o16 = torch.full((B * T, 2 * d), float("nan"), device=dev, dtype=torch.float16)
    ops.attn_fwd_tc(qkv16, in_splits=splits, B=B, T=T, H=H, probs16=probs16, o32=o32, o16=o16, o_splits=2, causal=causal)
    q, k, v = (t.reshape(B, T, H, 64).permute(0, 2, 1, 3) for t in qkv.view(B, T, 3 * d).chunk(3, -1))
    s = q @ k.transpose(-1, -2)
    if causal:
        s = s + torch.full((T, T), float("-inf"), device=dev).triu_(1)
    a = s.softmax(-1)
    o = (a @ v).permute(0, 2, 1, 3).reshape(B * T, d)
    tol = 1e-5 if splits == 2 else 2e-3
    assert torch.allclose(probs16[:, :, :T].float().view(B, H, T, T), a, atol=1e-3, rtol=1e-3)
    assert probs16[:, :, T:].abs().max().item() == 0.0 if ldp > T else True
    assert torch.allclose(o32, o, atol=tol, rtol=1e-4), (o32 - o).abs().max().item()
    assert torch.allclose(_recon(o16, d, 2), o, atol=tol, rtol=1e-4)
