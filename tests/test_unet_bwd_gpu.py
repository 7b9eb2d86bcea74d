"""Backward of the voxel UNet on the GPU (training step): every new kernel against torch fp32
autograd of the same op, then the whole ResidualUNet3D parameter / input gradients against autograd through the CPU
oracle (oracle/unet_oracle.py, pinned to the reference by oracle/gen_golden_3d.py).

Tolerance: data gradients run as 3-pass hi/lo fp16 MMAs (measured 5e-6 against the oracle on every intermediate);
weight gradients reduce single fp16 operands (per-tensor power-of-two scaling, fp32 accumulation), ~2^-11 relative
rounding per operand: per tensor ||Δ||_2 / ||ref||_2 <= 2e-3 (measured 3e-4 .. 7e-4)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
GRAD_TOL = 2e-3
KINK_TOL = 3e-2  # see _unet_grads_vs_oracle


def _rel2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _padded(x_cl, Cp=None):
    """[N,D,H,W,C] -> zero-padded fp16 [guard + N*(D+2)(H+2)(W+2) (rounded) + guard + 66, Cp] and the view at voxel 0"""
    N, D, H, W, C = x_cl.shape
    Cp = Cp or C
    PD, PH, PW = D + 2, H + 2, W + 2
    nv = N * PD * PH * PW
    guard = PH * PW + PW + 1
    nvr = (nv + 63) // 64 * 64
    buf = torch.zeros(guard + nvr + guard + 66, Cp, dtype=torch.float16, device=x_cl.device)
    return buf, buf[guard:], nvr, (PD, PH, PW)


@pytest.mark.parametrize("N,dims,Co,Ci", [(1, (4, 4, 8), 16, 16), (2, (8, 8, 8), 32, 32), (1, (4, 8, 16), 64, 32),
                                          (2, (4, 4, 4), 32, 16), (1, (4, 4, 4), 128, 128)])
def test_conv_wgrad_and_pack(N, dims, Co, Ci):
    from semabs_b200 import ops

    D, H, W = dims
    g = torch.Generator(device=dev).manual_seed(Co + Ci + D)
    x = torch.randn(N, D, H, W, Ci, device=dev, generator=g)
    dy = torch.randn(N, D, H, W, Co, device=dev, generator=g) * 3e-7   # far below fp16's normal range on purpose
    y = torch.randn(N, D, H, W, Co, device=dev, generator=g)           # ReLU mask source
    _, dz_pad, nvox, (PD, PH, PW) = _padded(dy)
    _, x_pad, _, _ = _padded(x)
    amax = torch.zeros(1, dtype=torch.int32, device=dev)
    scale = torch.zeros(1, device=dev)
    ops.absmax_f32(dy, amax)
    assert abs(amax.view(torch.float32).item() - dy.abs().max().item()) == 0
    dz_op = torch.empty(N, D, H, W, Co, dtype=torch.float16, device=dev)
    ops.unet_bwd_pack(dy, N=N, D=D, H=H, W=W, C=Co, amax=amax, mask=y, pad16=dz_pad, Cp=Co, op16=dz_op, scale_out=scale)
    ops.unet_bwd_pack(x, N=N, D=D, H=H, W=W, C=Ci, pad16=x_pad, Cp=Ci)
    s = scale.item()
    assert 2**12 <= dy.abs().max().item() * s < 2**13
    dz_ref = dy * (y > 0)
    assert _rel2(dz_op.float() / s, dz_ref) < 1e-3
    inner = dz_pad[: N * PD * PH * PW].view(N, PD, PH, PW, Co)[:, 1:-1, 1:-1, 1:-1]
    assert torch.equal(inner, dz_op)
    grad = torch.full((Co, Ci, 3, 3, 3), float("nan"), device=dev)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    seg_off = [(kd - 1) * PH * PW + (kh - 1) * PW - 1 for kd in range(3) for kh in range(3)]
    seg_slot = [[(s9 * 3) + kw for kw in range(3)] for s9 in range(9)]
    ops.conv3d_wgrad(dz_pad, x_pad, lda=Co, Ca=Co, Ca_real=Co, ldb=Ci, Cb=Ci, Cb_real=Ci, nvox=nvox, seg_off=seg_off,
                     seg_ntaps=[3] * 9, seg_sh=[[0, 1, 2]] * 9, seg_slot=seg_slot, nslots=27,
                     slot_k=torch.arange(27, dtype=torch.int32, device=dev), KT=27, workspace=ws, scale=scale, grad=grad)
    xr = x.permute(0, 4, 1, 2, 3).contiguous()
    w = torch.zeros(Co, Ci, 3, 3, 3, device=dev, requires_grad=True)
    F.conv3d(xr, w, padding=1).backward(dz_ref.permute(0, 4, 1, 2, 3).contiguous())
    assert _rel2(grad, w.grad) < GRAD_TOL, _rel2(grad, w.grad)


@pytest.mark.parametrize("N,dims,Ci,Co", [(1, (4, 4, 4), 32, 16), (2, (4, 4, 8), 64, 32)])
def test_conv_transpose_backward(N, dims, Ci, Co):
    """data gradient (conv kind 3, 8 chained parity launches) and weight gradient (parity-split dy) of
    ConvTranspose3d(k3, s2, p1, output_size = 2x)"""
    from semabs_b200 import ops

    D, H, W = dims
    g = torch.Generator(device=dev).manual_seed(Ci + D)
    x = torch.randn(N, Ci, D, H, W, device=dev, generator=g)
    w = (torch.randn(Ci, Co, 3, 3, 3, device=dev, generator=g) / (8 * Ci) ** 0.5).requires_grad_(True)
    dy = torch.randn(N, Co, 2 * D, 2 * H, 2 * W, device=dev, generator=g)
    xr = x.clone().requires_grad_(True)
    F.conv_transpose3d(xr, w, stride=2, padding=1, output_padding=1).backward(dy)
    dy_cl = dy.permute(0, 2, 3, 4, 1).contiguous()
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous()
    # pack: channels-last operand + 8 parity volumes
    PD, PH, PW = D + 2, H + 2, W + 2
    vol = N * PD * PH * PW
    guard = PH * PW + PW + 1
    nvr8 = (8 * vol + 63) // 64 * 64
    par_buf = torch.zeros(guard + nvr8 + guard + 66, Co, dtype=torch.float16, device=dev)
    dy_par = par_buf[guard:]
    _, x_pad, nvox, _ = _padded(x_cl)
    amax = torch.zeros(1, dtype=torch.int32, device=dev)
    scale = torch.zeros(1, device=dev)
    ops.absmax_f32(dy_cl, amax)
    dy_op = torch.empty(N, 2 * D, 2 * H, 2 * W, Co, dtype=torch.float16, device=dev)
    ops.unet_bwd_pack(dy_cl, N=N, D=2 * D, H=2 * H, W=2 * W, C=Co, amax=amax, pad16=dy_par, Cp=Co, parity=True, op16=dy_op,
                      scale_out=scale)
    ops.unet_bwd_pack(x_cl, N=N, D=D, H=H, W=W, C=Ci, pad16=x_pad, Cp=Ci)
    # data gradient
    wa = w.detach().permute(0, 2, 3, 4, 1).reshape(Ci, -1).half().contiguous()
    dx = torch.full((N, D, H, W, Ci), float("nan"), device=dev)
    for q in range(8):
        ops.conv3d(dy_op, wa, kind=ops.CONV_TRANSPOSE_ADJOINT, parity=q, N=N, D=D, H=H, W=W, C_in=Co, C_out=Ci, a_splits=1,
                   w_splits=1, precise=False, out32=dx, residual=dx if q > 0 else None)
    assert _rel2(dx.permute(0, 4, 1, 2, 3) / scale.item(), xr.grad) < GRAD_TOL
    # weight gradient
    par = lambda k: 0 if k == 1 else 1
    sh = lambda k: -1 if k == 0 else 0
    offA, offB = [], []
    for kz in range(3):
        for ky in range(3):
            base = sh(kz) * PH * PW + sh(ky) * PW
            offA.append(((par(kz) << 2) | (par(ky) << 1) | 1) * vol + base - 1)
            offB.append(((par(kz) << 2) | (par(ky) << 1) | 0) * vol + base)
    gw = torch.full((Ci, Co, 3, 3, 3), float("nan"), device=dev)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    kA = torch.tensor([(kz * 3 + ky) * 3 + kx for kz in range(3) for ky in range(3) for kx in (0, 2)], dtype=torch.int32, device=dev)
    kB = torch.tensor([(kz * 3 + ky) * 3 + 1 for kz in range(3) for ky in range(3)], dtype=torch.int32, device=dev)
    wk = dict(lda=Ci, Ca=Ci, Ca_real=Ci, ldb=Co, Cb=Co, Cb_real=Co, nvox=nvox, KT=27, workspace=ws, scale=scale, grad=gw)
    ops.conv3d_wgrad(x_pad, dy_par, seg_off=offA, seg_ntaps=[2] * 9, seg_sh=[[0, 1, 0]] * 9,
                     seg_slot=[[2 * s, 2 * s + 1, 0] for s in range(9)], nslots=18, slot_k=kA, **wk)
    ops.conv3d_wgrad(x_pad, dy_par, seg_off=offB, seg_ntaps=[1] * 9, seg_sh=[[0, 0, 0]] * 9,
                     seg_slot=[[s, 0, 0] for s in range(9)], nslots=9, slot_k=kB, **wk)
    assert _rel2(gw, w.grad) < GRAD_TOL, _rel2(gw, w.grad)


@pytest.mark.parametrize("N,S,C,C_real,groups", [(2, 512, 16, 16, 8), (1, 4096, 64, 64, 8), (2, 512, 16, 1, 1), (1, 64, 256, 256, 8)])
def test_groupnorm_backward(N, S, C, C_real, groups):
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(S + C)
    x = torch.randn(N, S, C, device=dev, generator=g) * 2 + 0.5
    x[..., C_real:] = 0
    dy = torch.randn(N, S, C, device=dev, generator=g)
    gamma = torch.zeros(C, device=dev)
    gamma[:C_real] = torch.randn(C_real, device=dev, generator=g)
    add = torch.randn(N, S, C, device=dev, generator=g)
    add_mask = torch.randn(N, S, C, device=dev, generator=g)
    prev = torch.randn(N, S, C, device=dev, generator=g)
    # reference through torch autograd (GroupNorm over the real channels)
    xr = x[..., :C_real].permute(0, 2, 1).contiguous().requires_grad_(True)
    gr = gamma[:C_real].clone().requires_grad_(True)
    br = torch.zeros(C_real, device=dev, requires_grad=True)
    F.group_norm(xr, groups, gr, br, eps=1e-5).backward(dy[..., :C_real].permute(0, 2, 1).contiguous())
    # ours
    cpg = C_real // groups if groups > 1 else C_real
    xg = x[..., :C_real].reshape(N, S, groups, cpg).double()
    stats = torch.stack([xg.sum(dim=(1, 3)), (xg**2).sum(dim=(1, 3))], dim=-1)
    stats8 = torch.zeros(N, 8, 2, dtype=torch.float64, device=dev)
    stats8[:, :groups] = stats
    stats_in = stats8[:, :groups].contiguous()
    sums = torch.zeros(N, C, 2, dtype=torch.float64, device=dev)
    scale = torch.tensor([4.0], device=dev)
    dys = dy * 4.0
    ops.groupnorm_bwd_reduce(dys, x, sums, N=N, S=S, C=C)
    dg = torch.empty(C_real, device=dev)
    db = torch.empty(C_real, device=dev)
    ops.groupnorm_param_grads(sums, N=N, S=S, C=C, C_real=C_real, groups=groups, stats=stats_in, scale=scale, dgamma=dg, dbeta=db)
    assert _rel2(dg, gr.grad) < 1e-4 and _rel2(db, br.grad) < 1e-4
    amax = torch.zeros(1, dtype=torch.int32, device=dev)
    dx = prev.clone()
    ops.groupnorm_bwd_apply(dys, x, stats_in, gamma, sums, dx, N=N, S=S, C=C, C_real=C_real, groups=groups, dy_scale=scale,
                            add=add * 2.0, add_scale=torch.tensor([2.0], device=dev), add_mask=add_mask, accumulate=True, amax=amax)
    ref = torch.zeros(N, S, C, device=dev)
    ref[..., :C_real] = xr.grad.permute(0, 2, 1)
    ref = ref + add * (add_mask > 0) + prev
    assert _rel2(dx, ref) < 1e-4, _rel2(dx, ref)
    assert abs(amax.view(torch.float32).item() - dx.abs().max().item()) < 1e-6 * dx.abs().max().item()


def test_maxpool_backward_first_max_rule():
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(3)
    N, D, C = 2, 8, 16
    x = torch.relu(torch.randn(N, C, D, D, D, device=dev, generator=g))  # many exact zeros -> ties
    x[:, :, :2] = 0
    gy = torch.randn(N, C, D // 2, D // 2, D // 2, device=dev, generator=g)
    xr = x.clone().requires_grad_(True)
    F.max_pool3d(xr, 2).backward(gy)
    prev = torch.randn(N, D, D, D, C, device=dev, generator=g)
    dx = prev.clone()
    amax = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.maxpool3d_2_bwd(gy.permute(0, 2, 3, 4, 1).contiguous() * 8, x.permute(0, 2, 3, 4, 1).contiguous(), dx, N=N, D=D, H=D,
                        W=D, C=C, g_scale=torch.tensor([8.0], device=dev), accumulate=True, amax=amax)
    ref = xr.grad.permute(0, 2, 3, 4, 1) + prev
    assert torch.allclose(dx, ref, rtol=1e-6, atol=1e-7)
    assert abs(amax.view(torch.float32).item() - dx.abs().max().item()) < 1e-6


def _unet_grads_vs_oracle(cin, cout, fmaps, levels, shape, N, seed, loss_scale=1.0, precise=True):
    from oracle import unet_oracle
    from semabs_b200.unet3d import ResidualUNet3D
    from tests._branches import branch_masks, oracle_on_our_branches, record_tapes

    torch.manual_seed(seed)
    m = ResidualUNet3D(in_channels=cin, out_channels=cout, f_maps=fmaps, num_groups=8, num_levels=levels, precise=precise).to(dev)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(N, cin, *shape, generator=g)
    gy = torch.randn(N, cout, *shape, generator=g) * loss_scale
    # precise=False: single fp16 operands in every convolution, forward and backward (the --use_amp counterpart): the
    # rounding is ~2^-11 per operand and layer instead of ~2^-22 (measured: worst tensor 2.2e-2, median 2.3e-3)
    tol, kink_tol, fwd_tol = (GRAD_TOL, KINK_TOL, 1e-3) if precise else (5e-2, 5e-2, 5e-3)
    # ours: training-mode forward (tape) + hand-written backward
    xg = x.to(dev).requires_grad_(True)
    with record_tapes() as tapes:
        y = m(xg)
    y.backward(gy.to(dev))
    # oracle: autograd through the CPU restatement of the reference module, on the ReLU branches our forward took
    # (tests/_branches.py explains why)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xo = x.clone().requires_grad_(True)
    with oracle_on_our_branches(branch_masks(tapes, levels)):
        yo = unet_oracle.residual_unet3d(sd, xo)
    yo.backward(gy)
    assert ((y.detach().cpu() - yo.detach()).abs().max() / yo.detach().abs().max()).item() < fwd_tol
    errs = {}
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        ref = sd[name].grad
        if ref is None or ref.norm() == 0:
            continue
        errs[name] = _rel2(p.grad.cpu(), ref)
    errs["<input>"] = _rel2(xg.grad.cpu(), xo.grad)
    ranked = sorted(errs.items(), key=lambda kv: -kv[1])
    print("gradient errors, worst first:", [(k, f"{v:.1e}") for k, v in ranked[:6]], "median", f"{ranked[len(ranked) // 2][1]:.1e}")
    # max-pool arg-max near-ties (two candidates within fp32 rounding) are the one remaining branch the oracle may take
    # differently; allow it to show in the <= 4 tensors of one unit
    over = [kv for kv in ranked if kv[1] >= tol]
    assert len(over) <= 4 and ranked[0][1] < kink_tol, ranked[:6]
    assert ranked[len(ranked) // 2][1] < (tol / 2 if precise else 1e-2)
    return ranked[0], errs["<input>"]


def test_unet_backward_matches_oracle_autograd():
    worst, ex = _unet_grads_vs_oracle(16, 16, 16, 3, (16, 16, 16), 2, seed=0)
    print("worst parameter gradient error", worst, "input gradient error", ex)


def test_unet_backward_tiny_gradients_and_padded_input_channels():
    # mean-BCE-sized upstream gradients (1e-8) and an input with fewer channels than one MMA K block / GroupNorm group
    _unet_grads_vs_oracle(1, 16, 16, 2, (8, 8, 8), 1, seed=2, loss_scale=1e-8)


def test_unet_backward_halo_level():
    # W = 128 takes the halo-resident conv kernel for the data gradients
    _unet_grads_vs_oracle(16, 16, 16, 2, (4, 8, 128), 1, seed=4)


def test_unet_backward_fast_mode():
    # precise=False end to end (what bench.py reports as the "amp_like" train step)
    _unet_grads_vs_oracle(16, 16, 16, 3, (16, 16, 16), 2, seed=6, precise=False)
    _unet_grads_vs_oracle(16, 16, 16, 2, (4, 8, 128), 1, seed=7, precise=False)
