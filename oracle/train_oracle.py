"""TEST INFRASTRUCTURE ONLY — CPU restatement of the optimiser side of the reference training step.
Pinned against the unmodified reference (`arm.optim.lamb.Lamb`, torch's binary_cross_entropy_with_logits /
clip_grad_norm_ as called by the reference) in oracle/gen_golden_3d.py:check_train_oracle."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def lamb_step(params, grads, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, adam=False):
    """arm/optim/lamb.py:59-127 on lists of tensors; `state` = list of dicts with exp_avg / exp_avg_sq (created on
    first use). Parameters whose grad is None are skipped (:71-72). In place."""
    for p, g, st in zip(params, grads, state):
        if g is None:
            continue
        if not st:
            st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
        m, v = st["exp_avg"], st["exp_avg_sq"]
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        weight_norm = p.pow(2).sum().sqrt().clamp(0, 10)
        adam_step = m / v.sqrt().add(eps)
        if weight_decay != 0:
            adam_step.add_(p, alpha=weight_decay)
        adam_norm = adam_step.pow(2).sum().sqrt()
        trust = 1 if (weight_norm == 0 or adam_norm == 0) else weight_norm / adam_norm
        if adam:
            trust = 1
        p.add_(adam_step, alpha=-lr * float(trust))


def clip_coefficient(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (utils.py:415): total L2 norm and the factor applied to every gradient."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads if g is not None)).float()
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def masked_bce(logits, labels, weight=None, ignore=None):
    """train_ovssc.py:133-150: BCE-with-logits over the non-ignored points, mean reduction; accuracy likewise."""
    keep = ~ignore if ignore is not None else torch.ones_like(logits, dtype=torch.bool)
    w = weight[keep] if weight is not None else None
    loss = F.binary_cross_entropy_with_logits(logits[keep], labels[keep].float(), weight=w)
    acc = ((logits > 0.0).long() == labels.long()).float()[keep].mean()
    return loss, acc
