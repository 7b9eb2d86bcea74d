"""Evaluation metrics on the device (semabs_b200.metrics) vs the CPU oracle (oracle/metrics_oracle.py, pinned
to the reference's utils.prediction_analysis / utils.voxelize_points). Counting work: results must be IDENTICAL."""
import numpy as np
import pytest
import torch

BOUNDS = ((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9))


def _case(seed, B, P, n):
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand(B, P, n, generator=g) < 0.3
    lab = torch.rand(B, P, n, generator=g) < 0.2
    ign = torch.rand(B, P, n, generator=g) < 0.25
    pred[0, 1] = False          # no predicted positives  -> precision NaN
    lab[B - 1, 0] = False       # no positive labels      -> recall NaN
    ign[B - 1, P - 1] = True    # everything ignored      -> every ratio NaN
    lo, hi = torch.tensor(BOUNDS[0]), torch.tensor(BOUNDS[1])
    xyz = lo + (hi - lo) * (torch.rand(B, P, n, 3, generator=g) * 1.1 - 0.05)  # some points outside the bounds (clamped)
    return pred, lab, ign, xyz


def test_oracle_known_answers():
    from oracle import metrics_oracle as mo

    pred = torch.tensor([[[1, 1, 0, 0, 1, 0]]]).bool()
    lab = torch.tensor([[[1, 0, 1, 0, 1, 1]]]).bool()
    ign = torch.tensor([[[0, 0, 0, 0, 0, 1]]]).bool()
    s = mo.prediction_analysis(pred, lab, ign)
    assert s["iou"] == [0.5] and s["precision"] == [2 / 3] and s["recall"] == [2 / 3]
    assert abs(s["false_negative"][0] - 0.2) < 1e-7 and abs(s["false_positive"][0] - 0.2) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("B,P,n,shape", [(2, 3, 5000, (8, 8, 8)), (1, 2, 40000, (32, 32, 32)), (2, 2, 777, (4, 6, 5))])
def test_metrics_match_oracle(B, P, n, shape):
    from oracle import metrics_oracle as mo
    from semabs_b200 import metrics

    pred, lab, ign, xyz = _case(B * 10 + P, B, P, n)
    a, b = metrics.prediction_analysis(pred, lab, ign), mo.prediction_analysis(pred, lab, ign)
    for k in b:
        assert np.allclose(a[k], b[k], equal_nan=True, rtol=5e-7, atol=0), k  # fp64 ratios of integers here, fp32 torch divisions / means in the reference
    va, vb = metrics.voxelize_points(pred, lab, xyz, shape, BOUNDS, ign), mo.voxelize_points(pred, lab, xyz, shape, BOUNDS, ign)
    for k in vb:
        assert torch.equal(va[k].cpu().float(), vb[k].float()), k
    sa = metrics.voxel_prediction_analysis(pred, lab, xyz, shape, BOUNDS, ign)
    sb = mo.prediction_analysis(vb["prediction"], vb["label"], vb["ignore"])
    for k in sb:
        assert np.allclose(sa[k], sb[k], equal_nan=True, rtol=5e-7, atol=0), k
