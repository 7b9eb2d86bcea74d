"""TEST INFRASTRUCTURE ONLY — golden outputs of the UNMODIFIED reference 3-D modules (unet3d.ResidualUNet3D,
net.SemAbs3D, net.SemAbsVOOL) on seeded weights/inputs -> tests/golden/unet_golden.npz, and the assertions that pin
oracle/unet_oracle.py to them.  Run here (needs /root/reference):  python -m oracle.gen_golden unet
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import ref_import, unet_oracle  # noqa: E402

BOUNDS = ((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9))


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max())


def semabs_args(voxel=32, levels=4, ch=16, **over):
    a = dict(voxel_shape=(voxel, voxel, voxel), scene_bounds=BOUNDS, unet_num_channels=ch, unet_f_maps=ch,
             unet_num_groups=8, unet_num_levels=levels, network_inputs=["saliency"], use_pts_feat_extractor=True,
             pts_feat_extractor_hidden_dim=128, reduce_method="max", device="cpu", batch_size=1)
    a.update(over)
    return a


def synth_points(seed, B, P, n_in, n_out):
    g = torch.Generator().manual_seed(seed)
    lo, hi = torch.tensor(BOUNDS[0]), torch.tensor(BOUNDS[1])
    xyz = lo + (hi - lo) * torch.rand(B, n_in, 3, generator=g)
    # a few points outside the bounds exercise the clamp
    xyz[:, :5] += 0.7
    feat = torch.randn(B, P, n_in, 1, generator=g)
    out_xyz = lo + (hi - lo) * (torch.rand(B, P, n_out, 3, generator=g) * 1.1 - 0.05)
    return xyz, feat, out_xyz


def main():
    out = {}
    unet3d = ref_import.import_reference_module("unet3d")
    net = ref_import.import_reference_module("net")

    # ---- ResidualUNet3D: 16^3, 3 levels (deepest 4^3), two samples --------------------------------
    torch.manual_seed(0)
    ref = unet3d.ResidualUNet3D(in_channels=16, out_channels=16, f_maps=16, num_groups=8, num_levels=3).eval()
    from semabs_b200.unet3d import ResidualUNet3D as Mine

    torch.manual_seed(0)
    mine = Mine(in_channels=16, out_channels=16, f_maps=16, num_groups=8, num_levels=3)
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys()), "state-dict key names/order differ from the reference"
    assert all(torch.equal(sd_ref[k], sd_mine[k]) for k in sd_ref), "same-seed initial values differ from the reference"
    x = torch.randn(2, 16, 16, 16, 16, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y_ref = ref(x)
        y_or = unet_oracle.residual_unet3d(sd_ref, x)
    print("UNet 16^3 oracle-vs-reference", rel_err(y_or, y_ref))
    assert rel_err(y_or, y_ref) < 1e-6
    out["unet16_out"] = y_ref.numpy()

    # ---- in_channels < num_groups (single group, channel padding) at 32^3, 4 levels ---------------
    torch.manual_seed(3)
    ref2 = unet3d.ResidualUNet3D(in_channels=2, out_channels=16, f_maps=32, num_groups=8, num_levels=4).eval()
    x2 = torch.randn(1, 2, 32, 32, 32, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        y2 = ref2(x2)
        assert rel_err(unet_oracle.residual_unet3d(ref2.state_dict(), x2), y2) < 1e-6
    out["unet32_c2_out_sub"] = y2[:, :, ::2, ::2, ::2].contiguous().numpy()

    # ---- SemAbs3D.forward, 32^3, 4 levels ------------------------------------------------------------
    torch.manual_seed(5)
    m = net.SemAbs3D(**semabs_args()).eval()
    assert m.vg.reduce_method == "mean"  # the load-bearing quirk
    xyz, feat, oxyz = synth_points(6, 1, 2, 2000, 3000)
    with torch.no_grad():
        logits = m(input_xyz_pts=xyz, input_feature_pts=feat, tsdf_vol=torch.ones(1, 1), output_xyz_pts=oxyz)
        lo = unet_oracle.semabs3d_forward(m.state_dict(), xyz, feat, oxyz, BOUNDS, (32, 32, 32))
    print("SemAbs3D oracle-vs-reference", rel_err(lo, logits), tuple(logits.shape))
    assert rel_err(lo, logits) < 1e-5
    out["semabs3d_logits"] = logits.numpy()
    from semabs_b200.net import SemAbs3D as MineNet

    torch.manual_seed(5)
    mm = MineNet(**semabs_args(device="cpu"))
    assert list(mm.state_dict().keys()) == list(m.state_dict().keys())
    assert all(torch.equal(mm.state_dict()[k], m.state_dict()[k]) for k in m.state_dict())

    # ---- SemAbsVOOL.forward ----------------------------------------------------------------------------
    torch.manual_seed(7)
    v = net.SemAbsVOOL(pointing_method="cosine_sim", pointing_dim=64, decoder_concat_xyz_pts=False, **semabs_args()).eval()
    g = torch.Generator().manual_seed(8)
    tgt, refsal = torch.randn(1, 2, 2000, 1, generator=g), torch.randn(1, 2, 2000, 1, generator=g)
    rel_names = [["behind"], ["on the left of"]]
    with torch.no_grad():
        vo = v(output_xyz_pts=oxyz, spatial_relation_name=rel_names, input_xyz_pts=xyz, input_target_saliency_pts=tgt,
               input_reference_saliency_pts=refsal, tsdf_vol=torch.ones(1, 1))
        voo = unet_oracle.semabsvool_forward(v.state_dict(), xyz, tgt, refsal, oxyz, rel_names, BOUNDS, (32, 32, 32))
    print("SemAbsVOOL oracle-vs-reference", rel_err(voo, vo), tuple(vo.shape))
    assert rel_err(voo, vo) < 1e-5
    out["vool_logits"] = vo.numpy()
    from semabs_b200.net import SemAbsVOOL as MineVool

    torch.manual_seed(7)
    mv = MineVool(pointing_method="cosine_sim", pointing_dim=64, decoder_concat_xyz_pts=False, **semabs_args(device="cpu"))
    assert list(mv.state_dict().keys()) == list(v.state_dict().keys())
    assert all(torch.equal(mv.state_dict()[k], v.state_dict()[k]) for k in v.state_dict())
    np.savez_compressed(os.path.join(GOLDEN, "unet_golden.npz"), **out)
    print("wrote unet_golden.npz")


def check_train_oracle():
    """Pins oracle/train_oracle.py against the reference's own Lamb (arm/optim/lamb.py) and torch's clip / BCE."""
    from oracle import train_oracle

    lamb_mod = ref_import.import_reference_module("arm.optim.lamb")
    g = torch.Generator().manual_seed(0)
    shapes = [(70000,), (33, 17), (5,), (8, 8)]
    ps = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    ps[2].data.zero_()  # weight_norm == 0 branch
    mine = [p.detach().clone() for p in ps]
    state = [dict() for _ in ps]
    opt = lamb_mod.Lamb(ps, lr=1e-2, betas=(0.9, 0.999), weight_decay=0.01)
    for step in range(3):
        grads = [torch.randn(*s, generator=g) * (10 if step == 1 else 1) for s in shapes]
        grads[3] = None  # a parameter without gradient is skipped
        for p, gr in zip(ps, grads):
            p.grad = None if gr is None else gr.clone()
        total_ref = torch.nn.utils.clip_grad_norm_(ps, 2.0)
        total, coef = train_oracle.clip_coefficient(grads, 2.0)
        assert torch.allclose(total, total_ref)
        opt.step()
        train_oracle.lamb_step(mine, [None if gr is None else gr * coef for gr in grads], state, lr=1e-2, weight_decay=0.01)
        for a, b in zip(mine, ps):
            assert torch.allclose(a, b.data, rtol=1e-6, atol=1e-7), (step, (a - b.data).abs().max())
    print("LAMB + clip oracle-vs-reference ok")


if __name__ == "__main__":
    main()
    check_train_oracle()
