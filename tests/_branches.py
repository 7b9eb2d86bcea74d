"""Test helper: make the oracle differentiate the SAME piecewise-linear branch as the CUDA forward took.

The UNet has ~10^6 ReLU inputs per test; our forward agrees with the oracle to ~5e-6, so every run a handful of
pre-activations within fp32 rounding of zero land on the other side of the kink (tools/debug_vool_kinks.py: |pre| ~ 1e-6
against an rms of 0.6).  Both gradients are then correct — for two different branches — and differ by O(1) in one element,
which a strongly cancelling reduction (GroupNorm bias gradient, or everything below a flipped top-level activation)
amplifies to 1e-2.  To compare like with like the oracle's backward is run with our branch decisions: its forward values
stay its own (relu(x)), only d relu / dx is taken from the masks recorded on our tapes."""
import contextlib

import torch
import torch.nn.functional as F


class _ReluWithBranch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return torch.relu(x)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


@contextlib.contextmanager
def record_tapes():
    """Collects the tapes of every training-mode UNet forward run inside the block (in call order)."""
    from semabs_b200 import unet3d_bwd

    tapes = []
    orig = unet3d_bwd.UNetBackward.new_tape

    def new_tape(self):
        t = orig(self)
        tapes.append(t)
        return t

    unet3d_bwd.UNetBackward.new_tape = new_tape
    try:
        yield tapes
    finally:
        unet3d_bwd.UNetBackward.new_tape = orig


def branch_masks(tapes, num_levels):
    """[ours > 0] for every ReLU of the reference forward, in the reference's call order: per UNet pass, blocks
    enc0..enc{L-1}, dec0..dec{L-2}; per block conv1's ReLU, conv2's ReLU, the residual ReLU (unet3d.py:243-259)."""
    masks = []
    blocks = [f"enc{i}" for i in range(num_levels)] + [f"dec{i}" for i in range(num_levels - 1)]
    for tape in tapes:
        n = tape.meta["N"]
        for b in blocks:
            rec = tape.blocks[b]
            for key in ("o1", "o2", "out"):
                t = rec[key].view(n, *rec["dims"], -1).permute(0, 4, 1, 2, 3)
                masks.append((t > 0).float().cpu())
    return masks


@contextlib.contextmanager
def oracle_on_our_branches(masks):
    it = iter(masks)
    orig = F.relu

    def relu(x, *a, **k):
        m = next(it)
        assert m.shape == x.shape, (m.shape, x.shape)
        return _ReluWithBranch.apply(x, m)

    F.relu = relu
    try:
        yield
    finally:
        F.relu = orig
    assert next(it, None) is None, "the oracle ran fewer ReLUs than the CUDA forward recorded"


@contextlib.contextmanager
def count_branch_flips(masks, result: dict):
    """Un-borrowed run: the oracle keeps its OWN ReLU branches; `result["n"]` receives the number of pre-activations whose
    sign decision differs from the masks recorded on our tapes (a wrong mask in the CUDA backward would show up here as a
    large count, fp32 rounding of the forward as a handful)."""
    it = iter(masks)
    orig = F.relu
    result["n"] = 0

    def relu(x, *a, **k):
        m = next(it)
        assert m.shape == x.shape, (m.shape, x.shape)
        result["n"] += int(((x.detach() > 0).float() != m).sum())
        return orig(x, *a, **k)

    F.relu = relu
    try:
        yield
    finally:
        F.relu = orig
    assert next(it, None) is None, "the oracle ran fewer ReLUs than the CUDA forward recorded"
