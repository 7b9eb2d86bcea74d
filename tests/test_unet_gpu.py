"""3-D side parity on the GPU: implicit-GEMM conv kernels vs torch fp32 convs, and the whole ResidualUNet3D /
SemAbs3D / SemAbsVOOL forward vs the committed outputs of the unmodified reference (tests/golden/unet_golden.npz,
made by oracle/gen_golden_3d.py) with seeded weights/inputs regenerated here."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"
# the torch-on-GPU comparisons below must be real fp32 (cuDNN convolutions default to TF32)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
GOLD = os.path.join(os.path.dirname(__file__), "golden", "unet_golden.npz")
BOUNDS = ((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9))
# spec tolerance (BASELINE.json north_star): voxel logits within 1e-3 relative = max|Δ| / max|ref|
TOL = 1e-3


@pytest.fixture(autouse=True)
def _inference_mode():
    # forward parity tests: under torch.no_grad() the modules take the inference path (workspace reuse, no tape);
    # the training path is covered by tests/test_unet_bwd_gpu.py and tests/test_train_gpu.py
    with torch.no_grad():
        yield


def _maxrel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def _cl(x, splits):  # NCDHW fp32 -> channels-last fp16 (hi|lo)
    xc = x.permute(0, 2, 3, 4, 1).contiguous()
    hi = xc.half()
    if splits == 1:
        return hi.contiguous()
    return torch.cat([hi, (xc - hi.float()).half()], dim=-1).contiguous()


def _pack(w2d, splits):
    hi = w2d.half()
    return hi.contiguous() if splits == 1 else torch.cat([hi, (w2d - hi.float()).half()], dim=1).contiguous()


@pytest.mark.parametrize("N,D,Ci,Co,precise", [(1, 16, 16, 16, True), (2, 8, 32, 64, True), (3, 4, 64, 128, False),
                                                (1, 32, 16, 32, False), (2, 4, 256, 256, True), (1, 8, 128, 16, True)])
def test_conv3x3x3(N, D, Ci, Co, precise):
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(N * 100 + D + Ci)
    s = 2 if precise else 1
    x = torch.randn(N, Ci, D, D, D, device=dev, generator=g)
    w = torch.randn(Co, Ci, 3, 3, 3, device=dev, generator=g) / (27 * Ci) ** 0.5
    res = torch.randn(N, D, D, D, Co, device=dev, generator=g)
    out = torch.full((N, D, D, D, Co), float("nan"), device=dev)
    out16 = torch.empty(N, D, D, D, s * Co, device=dev, dtype=torch.float16)
    G = 8
    stats = torch.zeros(N, G, 2, device=dev, dtype=torch.float64)
    ops.conv3d(_cl(x, s), _pack(w.permute(0, 2, 3, 4, 1).reshape(Co, -1), s), kind=ops.CONV_3X3X3, N=N, D=D, H=D, W=D,
               C_in=Ci, C_out=Co, a_splits=s, w_splits=s, precise=precise, residual=res, relu=True, out32=out, out16=out16,
               o16_splits=s, stats=stats, groups=G)
    torch.cuda.synchronize()
    xin = x if precise else x.half().float()
    win = w if precise else w.half().float()
    ref = F.relu(F.conv3d(xin, win, padding=1).permute(0, 2, 3, 4, 1) + res)
    tol = 5e-5 if precise else 1e-4  # precise: hi/lo split drops x_lo*w_lo (~2^-22 per term)
    assert _maxrel(out, ref) < tol, _maxrel(out, ref)
    got16 = out16[..., :Co].float() + (out16[..., Co:].float() if s == 2 else 0)
    assert _maxrel(got16, out) < (1e-5 if s == 2 else 2e-3)
    r = out.view(N, -1, G, Co // G)
    st_ref = torch.stack([r.double().sum(dim=(1, 3)), (r.double() ** 2).sum(dim=(1, 3))], dim=-1)
    assert torch.allclose(stats, st_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("N,D,Ci,Co", [(1, 4, 32, 16), (2, 8, 64, 32), (1, 16, 32, 16)])
def test_conv_transpose(N, D, Ci, Co):
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(D + Ci)
    x = torch.randn(N, Ci, D, D, D, device=dev, generator=g)
    w = torch.randn(Ci, Co, 3, 3, 3, device=dev, generator=g) / (8 * Ci) ** 0.5
    b = torch.randn(Co, device=dev, generator=g)
    skip = torch.randn(N, 2 * D, 2 * D, 2 * D, Co, device=dev, generator=g)
    out = torch.full((N, 2 * D, 2 * D, 2 * D, Co), float("nan"), device=dev)
    stats = torch.zeros(N, 8, 2, device=dev, dtype=torch.float64)
    wp = _pack(w.permute(1, 2, 3, 4, 0).reshape(Co, -1), 2)
    for parity in range(8):
        ops.conv3d(_cl(x, 2), wp, kind=ops.CONV_TRANSPOSE_PARITY, parity=parity, N=N, D=D, H=D, W=D, C_in=Ci, C_out=Co,
                   a_splits=2, w_splits=2, precise=True, bias=b, residual=skip, out32=out, stats=stats, groups=8)
    ref = F.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1).permute(0, 2, 3, 4, 1) + skip
    assert _maxrel(out, ref) < 2e-5, _maxrel(out, ref)
    r = ref.reshape(N, -1, 8, Co // 8)
    st_ref = torch.stack([r.double().sum(dim=(1, 3)), (r.double() ** 2).sum(dim=(1, 3))], dim=-1)
    assert torch.allclose(stats, st_ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("N,D,Ci,groups,Co", [(1, 8, 64, 8, 32), (2, 4, 128, 8, 32), (3, 4, 64, 1, 32), (1, 16, 64, 2, 32),
                                              (1, 8, 128, 8, 64), (2, 32, 64, 4, 64), (1, 4, 64, 2, 128)])
def test_conv_transpose_all_parities_one_launch(N, D, Ci, groups, Co):
    """semabs_conv_transpose3d_s2 (conv3d_convt.cu: eight output-parity classes per tile of 128 input voxels, weight slices of
    the classes sharing an input shift stacked along N) against torch's ConvTranspose3d and against the eight per-class launches."""
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(D + Ci + groups)
    x = torch.randn(N, Ci, D, D, D, device=dev, generator=g)
    w = torch.randn(Ci, Co, 3, 3, 3, device=dev, generator=g) / (8 * Ci) ** 0.5
    b = torch.randn(Co, device=dev, generator=g)
    skip = torch.randn(N, 2 * D, 2 * D, 2 * D, Co, device=dev, generator=g)
    wp = _pack(w.permute(1, 2, 3, 4, 0).reshape(Co, -1), 2)
    out = torch.full((N, 2 * D, 2 * D, 2 * D, Co), float("nan"), device=dev)
    stats = torch.zeros(N, groups, 2, device=dev, dtype=torch.float64)
    ops.conv_transpose3d_s2(_cl(x, 2), wp, N=N, D=D, H=D, W=D, C_in=Ci, C_out=Co, a_splits=2, w_splits=2, precise=True, bias=b,
                            residual=skip, out32=out, stats=stats, groups=groups)
    ref = F.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1).permute(0, 2, 3, 4, 1) + skip
    assert _maxrel(out, ref) < 2e-5, _maxrel(out, ref)
    r = ref.reshape(N, -1, groups, Co // groups)
    st_ref = torch.stack([r.double().sum(dim=(1, 3)), (r.double() ** 2).sum(dim=(1, 3))], dim=-1)
    assert torch.allclose(stats, st_ref, rtol=1e-5, atol=1e-5)
    out8 = torch.full_like(out, float("nan"))
    for parity in range(8):
        ops.conv3d(_cl(x, 2), wp, kind=ops.CONV_TRANSPOSE_PARITY, parity=parity, N=N, D=D, H=D, W=D, C_in=Ci, C_out=Co,
                   a_splits=2, w_splits=2, precise=True, bias=b, residual=skip, out32=out8)
    assert _maxrel(out, out8) < 5e-6, _maxrel(out, out8)  # same products, different fp32 accumulation order (measured 2.2e-6)
    # no bias / skip / statistics
    out2 = torch.full_like(out, float("nan"))
    ops.conv_transpose3d_s2(_cl(x, 2), wp, N=N, D=D, H=D, W=D, C_in=Ci, C_out=Co, a_splits=2, w_splits=2, precise=True, out32=out2)
    assert _maxrel(out2, ref - skip - b) < 2e-5


def test_groupnorm_pool_layout():
    from semabs_b200 import ops

    g = torch.Generator(device=dev).manual_seed(11)
    N, C, D = 2, 32, 8
    x = torch.randn(N, C, D, D, D, device=dev, generator=g) * 2 + 1
    S = D**3
    raw = torch.empty(N, S, C, device=dev)
    st = torch.zeros(N, 8, 2, device=dev, dtype=torch.float64)
    ops.ncdhw_to_ndhwc(x, raw, N=N, S=S, C=C, Cpad=C, groups=8, stats=st)
    assert torch.equal(raw.view(N, D, D, D, C), x.permute(0, 2, 3, 4, 1))
    gamma, beta = torch.randn(C, device=dev, generator=g), torch.randn(C, device=dev, generator=g)
    y16 = torch.empty(N, S, 2 * C, device=dev, dtype=torch.float16)
    ops.groupnorm_apply(raw, st, gamma, beta, y16, N=N, S=S, C=C, C_real=C, groups=8, splits=2)
    ref = F.group_norm(x, 8, gamma, beta, 1e-5).permute(0, 2, 3, 4, 1).reshape(N, S, C)
    got = y16[..., :C].float() + y16[..., C:].float()
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5)
    pooled = torch.empty(N, S // 8, C, device=dev)
    pst = torch.zeros(N, 8, 2, device=dev, dtype=torch.float64)
    ops.maxpool3d_2(raw, pooled, N=N, D=D, H=D, W=D, C=C, groups=8, stats=pst)
    pref = F.max_pool3d(x, 2)
    assert torch.equal(pooled.view(N, D // 2, D // 2, D // 2, C), pref.permute(0, 2, 3, 4, 1))
    r = pref.reshape(N, 8, -1).double()
    assert torch.allclose(pst, torch.stack([r.sum(-1), (r**2).sum(-1)], -1), rtol=1e-5)
    back = torch.empty(N, C, D, D, D, device=dev)
    ops.ndhwc_to_ncdhw(raw, back, N=N, S=S, C=C)
    assert torch.equal(back, x)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("precise", [True, False])
def test_unet_matches_reference(gold, precise):
    from semabs_b200.unet3d import ResidualUNet3D

    torch.manual_seed(0)
    m = ResidualUNet3D(in_channels=16, out_channels=16, f_maps=16, num_groups=8, num_levels=3, precise=precise).to(dev)
    x = torch.randn(2, 16, 16, 16, 16, generator=torch.Generator().manual_seed(1)).to(dev)
    y = m(x)
    ref = torch.from_numpy(gold["unet16_out"]).to(dev)
    err = _maxrel(y, ref)
    print(f"UNet 16^3 precise={precise}: max|d|/max|ref| = {err:.2e}")
    assert err < (TOL if precise else 2e-2)
    if precise:
        assert (y.argmax(1) == ref.argmax(1)).float().mean().item() > 0.999


def test_unet_few_input_channels(gold):
    from semabs_b200.unet3d import ResidualUNet3D

    torch.manual_seed(3)
    m = ResidualUNet3D(in_channels=2, out_channels=16, f_maps=32, num_groups=8, num_levels=4).to(dev)
    x = torch.randn(1, 2, 32, 32, 32, generator=torch.Generator().manual_seed(4)).to(dev)
    y = m(x)[:, :, ::2, ::2, ::2]
    ref = torch.from_numpy(gold["unet32_c2_out_sub"]).to(dev)
    err = _maxrel(y, ref)
    print(f"UNet 32^3 in_channels=2: {err:.2e}")
    assert err < TOL


def _semabs_args(**over):
    a = dict(voxel_shape=(32, 32, 32), scene_bounds=BOUNDS, unet_num_channels=16, unet_f_maps=16, unet_num_groups=8,
             unet_num_levels=4, network_inputs=["saliency"], use_pts_feat_extractor=True,
             pts_feat_extractor_hidden_dim=128, reduce_method="max", device="cuda", batch_size=1)
    a.update(over)
    return a


def _points(seed, B, P, n_in, n_out):
    g = torch.Generator().manual_seed(seed)
    lo, hi = torch.tensor(BOUNDS[0]), torch.tensor(BOUNDS[1])
    xyz = lo + (hi - lo) * torch.rand(B, n_in, 3, generator=g)
    xyz[:, :5] += 0.7
    feat = torch.randn(B, P, n_in, 1, generator=g)
    out_xyz = lo + (hi - lo) * (torch.rand(B, P, n_out, 3, generator=g) * 1.1 - 0.05)
    return xyz, feat, out_xyz


def test_semabs3d_matches_reference(gold):
    from semabs_b200.net import SemAbs3D

    torch.manual_seed(5)
    m = SemAbs3D(**_semabs_args()).to(dev)
    xyz, feat, oxyz = _points(6, 1, 2, 2000, 3000)
    logits = m(input_xyz_pts=xyz.to(dev), input_feature_pts=feat.to(dev), tsdf_vol=torch.ones(1, 1, device=dev),
               output_xyz_pts=oxyz.to(dev))
    ref = torch.from_numpy(gold["semabs3d_logits"]).to(dev)
    err = _maxrel(logits, ref)
    print(f"SemAbs3D logits: {err:.2e}")
    assert logits.shape == ref.shape and err < TOL
    # argmax over patches ("voxel labels") must agree wherever the reference's margin is not a numerical tie
    margin = (ref[:, 0] - ref[:, 1]).abs() > 4 * TOL * ref.abs().max()
    assert ((logits.argmax(1) == ref.argmax(1)) | ~margin).all()
    vol = m.visual_volumetric_features
    assert vol.shape == (2, 16, 32, 32, 32)


def test_semabsvool_matches_reference(gold):
    from semabs_b200.net import SemAbsVOOL

    torch.manual_seed(7)
    v = SemAbsVOOL(pointing_method="cosine_sim", pointing_dim=64, decoder_concat_xyz_pts=False, **_semabs_args()).to(dev)
    xyz, _, oxyz = _points(6, 1, 2, 2000, 3000)
    g = torch.Generator().manual_seed(8)
    tgt, refsal = torch.randn(1, 2, 2000, 1, generator=g), torch.randn(1, 2, 2000, 1, generator=g)
    out = v(output_xyz_pts=oxyz.to(dev), spatial_relation_name=[["behind"], ["on the left of"]], input_xyz_pts=xyz.to(dev),
            input_target_saliency_pts=tgt.to(dev), input_reference_saliency_pts=refsal.to(dev), tsdf_vol=torch.ones(1, 1, device=dev))
    ref = torch.from_numpy(gold["vool_logits"]).to(dev)
    err = _maxrel(out, ref)
    print(f"SemAbsVOOL logits: {err:.2e}")
    assert out.shape == ref.shape and err < TOL


@pytest.mark.parametrize("pair", [True, False])
@pytest.mark.parametrize("N,D,H,Ci,Co,precise", [(1, 4, 6, 16, 16, True), (2, 8, 8, 32, 32, True), (1, 16, 5, 32, 16, False),
                                                  (1, 32, 4, 16, 32, True), (3, 6, 10, 32, 32, False), (1, 4, 5, 32, 32, True),
                                                  (2, 64, 12, 32, 32, True), (1, 8, 6, 16, 32, False)])
def test_conv_halo_resident(N, D, H, Ci, Co, precise, pair, request):
    """semabs_conv3d_halo (W = 128 level): chunk-planar input via groupnorm_apply(planar), per-tap weight images.
    pair=True: C_out = 32 shapes with an even item count take the CTA-pair kernel (conv3d_halo2.cu, cta_group::2 MMAs); the
    (1, 4, 5, ...) case has an odd item count and exercises the fall-back to the single-CTA kernel."""
    from semabs_b200 import ops

    ops.set_halo_pair(pair)
    request.addfinalizer(lambda: ops.set_halo_pair(True))

    W = 128
    g = torch.Generator(device=dev).manual_seed(D * 10 + Ci + Co)
    s = 2 if precise else 1
    S = D * H * W
    x = torch.randn(N, Ci, D, H, W, device=dev, generator=g) * 1.5 + 0.3
    w = torch.randn(Co, Ci, 3, 3, 3, device=dev, generator=g) / (27 * Ci) ** 0.5
    gamma, beta = 1 + 0.2 * torch.randn(Ci, device=dev, generator=g), 0.2 * torch.randn(Ci, device=dev, generator=g)
    G = 8
    raw = torch.empty(N, S, Ci, device=dev)
    st = torch.zeros(N, G, 2, device=dev, dtype=torch.float64)
    ops.ncdhw_to_ndhwc(x, raw, N=N, S=S, C=Ci, Cpad=Ci, groups=G, stats=st)
    xn = torch.empty(N * s * Ci * S, device=dev, dtype=torch.float16)
    ops.groupnorm_apply(raw, st, gamma, beta, xn, N=N, S=S, C=Ci, C_real=Ci, groups=G, splits=s, planar=True)
    res = torch.randn(N, D, H, W, Co, device=dev, generator=g)
    out = torch.full((N, D, H, W, Co), float("nan"), device=dev)
    out16 = torch.empty(N, D, H, W, s * Co, device=dev, dtype=torch.float16)
    stats = torch.zeros(N, G, 2, device=dev, dtype=torch.float64)
    ops.conv3d_halo(xn, ops.pack_halo_weights(w, s), N=N, D=D, H=H, W=W, C_in=Ci, C_out=Co, a_splits=s, w_splits=s,
                    precise=precise, residual=res, relu=True, out32=out, out16=out16, o16_splits=s, stats=stats, groups=G)
    torch.cuda.synchronize()
    xg = F.group_norm(x, G, gamma, beta, 1e-5)
    if not precise:
        xg, w = xg.half().float(), w.half().float()
    ref = F.relu(F.conv3d(xg, w, padding=1).permute(0, 2, 3, 4, 1) + res)
    err = _maxrel(out, ref)
    assert err < (5e-5 if precise else 2e-4), err
    got16 = out16[..., :Co].float() + (out16[..., Co:].float() if s == 2 else 0)
    assert _maxrel(got16, out) < (1e-5 if s == 2 else 2e-3)
    r = out.view(N, -1, G, Co // G)
    st_ref = torch.stack([r.double().sum(dim=(1, 3)), (r.double() ** 2).sum(dim=(1, 3))], dim=-1)
    assert torch.allclose(stats, st_ref, rtol=1e-5, atol=1e-6)


def test_unet_full_resolution_row_path():
    """A UNet whose first level is 128 voxels wide goes through the halo-resident kernel; compare with the oracle."""
    from oracle import unet_oracle
    from semabs_b200.unet3d import ResidualUNet3D

    torch.manual_seed(11)
    m = ResidualUNet3D(in_channels=16, out_channels=16, f_maps=16, num_groups=8, num_levels=3).to(dev)
    x = torch.randn(1, 16, 16, 16, 128, generator=torch.Generator().manual_seed(12))
    y = m(x.to(dev)).cpu()
    with torch.no_grad():
        ref = unet_oracle.residual_unet3d({k: v.cpu() for k, v in m.state_dict().items()}, x)
    err = _maxrel(y, ref)
    print(f"UNet 16x16x128 (halo path): {err:.2e}")
    assert err < TOL
    m.use_halo = False
    y2 = m(x.to(dev)).cpu()
    assert _maxrel(y2, ref) < TOL


@pytest.mark.parametrize("shape", [(1, 16, 16, 128), (2, 8, 32, 128)])
def test_unet_folded_groupnorm_matches_unfused_and_oracle(shape):
    """128-wide level with 32 channels, inference: conv2 / conv3 of every block read the producer's RAW planar output with
    the GroupNorm folded into per-sample weights + a 27-class border bias (semabs_conv3d_halo_fused).  Against the oracle and
    against the un-folded path (GroupNorm-apply passes) of the same module."""
    from oracle import unet_oracle
    from semabs_b200.unet3d import ResidualUNet3D

    N, D, H, W = shape
    torch.manual_seed(13)
    m = ResidualUNet3D(in_channels=32, out_channels=32, f_maps=32, num_groups=8, num_levels=3).to(dev)
    x = torch.randn(N, 32, D, H, W, generator=torch.Generator().manual_seed(14)) * 1.5 + 0.2
    m.fold_groupnorm = True
    y = m(x.to(dev)).cpu()
    assert m.folded_blocks == 2, "the folded path was not taken (first encoder and last decoder block)"
    with torch.no_grad():
        ref = unet_oracle.residual_unet3d({k: v.cpu() for k, v in m.state_dict().items()}, x)
    err = _maxrel(y, ref)
    m.fold_groupnorm = False
    y2 = m(x.to(dev)).cpu()
    assert m.folded_blocks == 2
    err2, diff = _maxrel(y2, ref), _maxrel(y, y2)
    print(f"UNet {shape} folded GroupNorm: vs oracle {err:.2e} (un-folded path {err2:.2e}), folded vs un-folded {diff:.2e}")
    assert err < TOL and err2 < TOL and diff < 1e-4


def test_unet_cuda_graph_replay_matches_eager():
    """Inference forward: first call eager, second call with the same input buffer captures a CUDA graph, later calls replay it.
    The replay must follow in-place changes of the input, count its launches, and be dropped when a workspace is replaced."""
    from semabs_b200.unet3d import ResidualUNet3D

    torch.manual_seed(21)
    m = ResidualUNet3D(in_channels=32, out_channels=32, f_maps=32, num_groups=8, num_levels=3).to(dev)
    x = torch.randn(2, 32, 8, 16, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(22))
    m.use_cuda_graph = False
    ref = m(x).clone()
    m.use_cuda_graph = True
    y1 = m(x)  # eager (first sight of this buffer)
    assert not m._graphs
    y2 = m(x)  # capture + replay
    assert len(m._graphs) == 1
    l0 = m.kernel_launches
    y3 = m(x)  # replay
    assert m.kernel_launches - l0 > 10, "replays must keep counting the launches they stand for"
    for y in (y1, y2, y3):
        assert _maxrel(y, ref) < 5e-6
    assert y3.data_ptr() != y2.data_ptr()
    # same buffer, new values (not an affine change: the first GroupNorm would undo it)
    x.copy_(torch.randn(x.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(23)))
    m.use_cuda_graph = False
    ref2 = m(x).clone()
    m.use_cuda_graph = True
    assert _maxrel(m(x), ref2) < 5e-6 and _maxrel(ref2, ref) > 1e-2
    xs = x[:1].contiguous()  # another batch size replaces workspaces: every graph is dropped, results stay right
    m.use_cuda_graph = False
    ref3 = m(xs).clone()
    m.use_cuda_graph = True
    assert _maxrel(m(xs), ref3) < 5e-6 and not m._graphs
    assert _maxrel(m(x), ref2) < 5e-6
    assert _maxrel(m(x), ref2) < 5e-6 and len(m._graphs) == 1
    with torch.no_grad():
        m.final_conv.bias.add_(1.0)  # a changed parameter: new weight pack, new graph key
    m.use_cuda_graph = False
    ref4 = m(x).clone()
    m.use_cuda_graph = True
    m(x)
    assert _maxrel(m(x), ref4) < 5e-6 and _maxrel(ref4, ref2) > 1e-2
