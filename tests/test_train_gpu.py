"""Optimiser side of the training step on the GPU (BCE loss/gradient, clip_grad_norm_, LAMB) vs the oracle
restatement of the reference (oracle/train_oracle.py, pinned to arm/optim/lamb.py by oracle/gen_golden_3d.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def test_masked_bce_loss_and_gradient():
    from oracle import train_oracle
    from semabs_b200 import train

    g = torch.Generator().manual_seed(0)
    x = (torch.randn(2, 3, 5000, generator=g) * 3).requires_grad_(True)
    y = (torch.rand(2, 3, 5000, generator=g) < 0.1).float()
    w = torch.rand(2, 3, 5000, generator=g) + 0.5
    ig = torch.rand(2, 3, 5000, generator=g) < 0.3
    ig[1, 2] = True  # a fully padded patch
    loss_ref, acc_ref = train_oracle.masked_bce(x, y, w, ig)
    loss_ref.backward()
    loss, acc, dx = train.bce_with_logits_masked(x.detach().to(dev), y.to(dev), w.to(dev), ig.to(dev))
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * abs(loss_ref.item())
    assert abs(acc.item() - acc_ref.item()) < 1e-6
    assert torch.allclose(dx.cpu(), x.grad, rtol=1e-4, atol=1e-9)
    # no weights / no mask
    l2, _, _ = train.bce_with_logits_masked(x.detach().to(dev), y.to(dev))
    assert abs(l2.item() - train_oracle.masked_bce(x.detach(), y)[0].item()) < 1e-5


def test_lamb_with_clipping_matches_reference_semantics():
    from oracle import train_oracle
    from semabs_b200 import train

    g = torch.Generator().manual_seed(1)
    shapes = [(70000,), (33, 17), (5,), (8, 8), (200000,)]
    init = [torch.randn(*s, generator=g) for s in shapes]
    init[2].zero_()  # weight_norm == 0 -> trust ratio 1
    cpu = [t.clone() for t in init]
    state = [dict() for _ in cpu]
    params = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
    opt = train.Lamb(params, lr=1e-2, weight_decay=0.01)
    for step in range(3):
        grads = [torch.randn(*s, generator=g) * (10 if step == 1 else 0.001) for s in shapes]
        grads[3] = None  # unused parameter: skipped by both
        for p, gr in zip(params, grads):
            p.grad = None if gr is None else gr.clone().to(dev)
        total, coef = train_oracle.clip_coefficient(grads, 2.0)
        train_oracle.lamb_step(cpu, [None if gr is None else gr * coef for gr in grads], state, lr=1e-2, weight_decay=0.01)
        opt.step(max_grad_norm=2.0)
        for a, p in zip(cpu, params):
            assert torch.allclose(a, p.data.cpu(), rtol=2e-5, atol=1e-6), (step, (a - p.data.cpu()).abs().max())
    assert torch.equal(params[3].data.cpu(), init[3])
    # stand-alone clip_grad_norm_ drop-in
    for p in params:
        p.grad = None if p is params[3] else torch.ones_like(p) * 0.5
    tn = train.clip_grad_norm_(params, 1.0)
    ref_total = (sum(p.numel() for p in params if p.grad is not None) * 0.25) ** 0.5
    assert abs(tn.item() - ref_total) < 1e-3 * ref_total
    assert abs(params[0].grad[0].item() - 0.5 / (ref_total + 1e-6)) < 1e-6
