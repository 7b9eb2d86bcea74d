"""GPU suite for the training-side boundary: `train_ovssc.get_losses` / `train_vool.get_losses` (stats + per-cutoff
DataFrame) against the outputs of the UNMODIFIED reference functions (tests/golden/train_golden.json, made by
oracle/gen_golden_train.py), and the checkpoint round trip of `utils.get_net` / `utils.loop` / `utils.train`."""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "train_golden.json")))


def _check_frame(df, gold):
    assert list(df.columns) == gold["columns"]
    rows = list(df.itertuples(index=False, name=None))
    assert len(rows) == len(gold["rows"])
    n_num = n_nan = 0
    for r, g in zip(rows, gold["rows"]):
        for col, a, b in zip(gold["columns"], r, g):
            if isinstance(b, str):
                assert a == b, (col, a, b)
            elif b is None:
                assert isinstance(a, float) and math.isnan(a), (col, a)
                n_nan += 1
            else:
                assert abs(float(a) - b) <= 1e-6 * max(1.0, abs(b)), (col, a, b)
                n_num += 1
    return n_num, n_nan


@pytest.mark.parametrize("bal", [False, True])
def test_train_ovssc_get_losses_matches_reference(bal):
    from oracle.gen_golden_train import CUTOFFS_OVSSC, StubNet, make_ovssc_case
    from semabs_b200 import train_ovssc

    logits, batch = make_ovssc_case()
    sb = batch.pop("scene_bounds")
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    x = logits.to(dev).requires_grad_(True)
    stats, df = train_ovssc.get_losses(StubNet(x, dev), batch, cutoffs=CUTOFFS_OVSSC, balance_positive_negative=bal, scene_bounds=sb)
    gold = GOLD[f"ovssc_bal{int(bal)}"]
    assert set(stats) == set(gold["stats"])
    for k, v in gold["stats"].items():
        assert abs(float(stats[k]) - v) < 1e-5 * max(1.0, abs(v)), (k, float(stats[k]), v)
    n_num, n_nan = _check_frame(df, gold["frame"])
    stats["loss"].backward()
    assert abs(float(x.grad.abs().sum()) - gold["dlogits_sum_abs"]) < 1e-4 * gold["dlogits_sum_abs"]
    assert np.allclose(x.grad.flatten()[:8].cpu().numpy(), gold["dlogits_head"], rtol=1e-4, atol=1e-10)
    print(f"train_ovssc.get_losses(balance={bal}): {n_num} numeric cells and {n_nan} NaN cells identical to the reference's DataFrame")


@pytest.mark.parametrize("bal", [False, True])
def test_train_vool_get_losses_matches_reference(bal):
    from oracle.gen_golden_train import CUTOFFS_VOOL, StubNet, make_vool_case
    from semabs_b200 import train_vool

    logits, batch = make_vool_case()
    sb = batch.pop("scene_bounds")
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    x = logits.to(dev).requires_grad_(True)
    stats, df = train_vool.get_losses(StubNet(x, dev), batch, cutoffs=CUTOFFS_VOOL, balance_positive_negative=bal, scene_bounds=sb)
    gold = GOLD[f"vool_bal{int(bal)}"]
    assert set(stats) == set(gold["stats"])
    for k, v in gold["stats"].items():
        assert abs(float(stats[k]) - v) < 1e-5 * max(1.0, abs(v)), (k, float(stats[k]), v)
    _check_frame(df, gold["frame"])
    stats["loss"].backward()
    assert abs(float(x.grad.abs().sum()) - gold["dlogits_sum_abs"]) < 1e-4 * gold["dlogits_sum_abs"]


def _args(tmp, extra=()):
    from semabs_b200 import utils

    a = utils.config_parser().parse_args(["--file_path", "synthetic", "--voxel_shape", "32", "--unet_num_levels", "4",
                                          "--num_input_pts", "3000", "--num_output_pts", "5000", "--num_workers", "0",
                                          "--num_warmup_steps", "2", "--epochs", "1", "--lr", "0.01", *extra])
    a.voxel_shape = [32, 32, 32]
    a.log = str(tmp)
    return a


@pytest.mark.parametrize("task", ["ovssc", "vool"])
def test_train_save_reload_continue(tmp_path, task):
    """3 optimiser steps through utils.train (LAMB + cosine_with_restarts scheduler + checkpoint), reload the checkpoint
    into a FRESH network / optimiser through utils.get_net(load=...), and take step 4 on both: the restored state is
    bit-identical to the live one, and step 4 agrees.  (Step 4's loss itself is not bit-reproducible run to run — the
    point scatter and the GroupNorm statistics use floating-point atomics — so it is compared to 1e-5, while every
    restored tensor is compared bit for bit.)"""
    from torch.utils.data import DataLoader

    from semabs_b200 import train_ovssc, train_vool, utils
    from semabs_b200.net import SemAbs3D, SemAbsVOOL

    mod, net_class, ds_class = (train_ovssc, SemAbs3D, utils.SyntheticOVSSCDataset) if task == "ovssc" else \
        (train_vool, SemAbsVOOL, utils.SyntheticVOOLDataset)
    args = _args(tmp_path)
    exp = utils.setup_experiment(args=args, net_class=net_class, dataset_class=ds_class, length=3)
    net, opt, sched = exp["net"], exp["optimizer"], exp["lr_scheduler"]
    utils.train(get_losses_fn=mod.get_losses, **exp, **vars(args))
    assert float(net.steps) == 3.0
    ck = torch.load(os.path.join(args.log, "latest.pth"), map_location=dev, weights_only=False)
    assert ck["epochs"] == 1 and set(ck) == {"net", "optimizer", "epochs"}
    assert os.path.exists(os.path.join(args.log, "ckpt_0.pth")) and os.path.exists(os.path.join(args.log, "detailed_stats.pkl"))
    # a DDP-style checkpoint (reference: keys carry "module.") must load as well
    ck_ddp = dict(ck, net={"module." + k: v for k, v in ck["net"].items()})
    torch.save(ck_ddp, os.path.join(args.log, "ddp.pth"))
    args2 = _args(tmp_path)
    args2.load = os.path.join(args.log, "ddp.pth")
    net2, opt2, sched2, start_epoch, scaler = utils.get_net(train_dataset=exp["datasets"]["train"], net_class=net_class, **vars(args2))
    assert start_epoch == 1 and scaler is None
    for (k, a), (_, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a, b), k
    for p, q in zip(net.parameters(), net2.parameters()):
        if p in opt.state:
            assert opt2.state[q]["step"] == opt.state[p]["step"] <= 3  # (a relation embedding absent from a batch skips that step)
            assert torch.equal(opt.state[p]["exp_avg"], opt2.state[q]["exp_avg"]) and torch.equal(opt.state[p]["exp_avg_sq"], opt2.state[q]["exp_avg_sq"])
        else:
            assert q not in opt2.state or len(opt2.state[q]) == 0  # unused parameters never got optimiser state
    assert max(st["step"] for st in opt.state.values() if st) == 3
    # step 4 on both (same batch, same learning rate)
    for g in opt2.param_groups:
        g["lr"] = opt.param_groups[0]["lr"]
    batch = next(iter(DataLoader(exp["datasets"]["train"], batch_size=1)))
    out = []
    for n_, o_ in ((net, opt), (net2, opt2)):
        df = utils.loop(net=n_, loader=[dict(batch)], pbar=None, get_losses_fn=mod.get_losses, optimizer=o_, lr_scheduler=None,
                        grad_max_norm=2.0, device=torch.device(dev), scene_bounds=args.scene_bounds, cutoffs=[-1.0])
        assert len(df) == (args.num_patches if task == "ovssc" else args.num_descs)
        out.append({k: v.clone() for k, v in n_.state_dict().items()})
    worst = max(((out[0][k].float() - out[1][k].float()).abs().max() / out[0][k].float().abs().max().clamp_min(1e-12)).item() for k in out[0])
    print(f"{task}: after step 4, continued vs reloaded parameters differ by at most {worst:.1e} (relative)")
    assert worst < 1e-4 and float(net2.steps) == 4.0
