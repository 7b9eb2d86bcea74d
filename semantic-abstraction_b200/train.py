"""Optimiser side of the reference training step on the B200 kernels: masked BCE-with-logits (loss + gradient),
gradient-norm clipping and LAMB — mirrors of `train_ovssc.get_losses`' loss (train_ovssc.py:133-150),
`torch.nn.utils.clip_grad_norm_` as used at utils.py:415 and `arm.optim.lamb.Lamb` (arm/optim/lamb.py:25-127),
plus the data-parallel gradient all-reduce (the one real exchange step of the path, utils.py:255-258).

The gradients themselves come from the hand-written backward of SemAbs3D / SemAbsVOOL / ResidualUNet3D (net.py
autograd nodes -> unet3d_bwd.py, csrc/unet_bwd.cu, csrc/points_bwd.cu).  `get_losses_ovssc` / `get_losses_vool` mirror the
loss part of the reference's `get_losses` (train_ovssc.py:81-150, train_vool.py:118-185; the IoU tables are evaluation,
SURVEY.md §8f), `train_step` mirrors the train branch of `utils.loop` (utils.py:404-422).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from ._lib import check, f32, i32, lib, ptr, stream_ptr
from .unet3d import bump_weights_epoch

CHUNK = 65536


def bce_with_logits_masked(logits: torch.Tensor, labels: torch.Tensor, weight: Optional[torch.Tensor] = None,
                           ignore: Optional[torch.Tensor] = None, need_grad: bool = True):
    """-> (loss [scalar tensor], accuracy [scalar tensor], dlogits or None); all on the device of `logits`."""
    assert logits.is_cuda and logits.dtype == torch.float32
    x = logits.contiguous()
    y = labels.to(torch.float32).contiguous()
    w = weight.to(torch.float32).contiguous() if weight is not None else None
    ig = ignore.to(torch.uint8).contiguous() if ignore is not None else None
    acc = torch.empty(3, dtype=torch.float64, device=x.device)
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x) if need_grad else None
    check(lib().semabs_bce_with_logits(ptr(x), ptr(y), ptr(w), ptr(ig), C.c_int64(x.numel()), ptr(acc), ptr(out), ptr(dx),
                                       stream_ptr()))
    return out[0], out[1], dx


class _BceFn(torch.autograd.Function):
    """binary_cross_entropy_with_logits(outputs[keep], labels[keep], weight[keep]) (mean) as one kernel; the same launch
    produces the gradient, backward only scales it."""

    @staticmethod
    def forward(ctx, logits, labels, weight, ignore):
        loss, acc, dx = bce_with_logits_masked(logits, labels, weight, ignore, need_grad=True)
        ctx.save_for_backward(dx)
        ctx.mark_non_differentiable(acc)
        return loss.clone(), acc.clone()

    @staticmethod
    def backward(ctx, g_loss, g_acc):
        (dx,) = ctx.saved_tensors
        return dx * g_loss, None, None, None


def get_bce_weight(output_label_pts: torch.Tensor, balance_positive_negative: bool) -> Optional[torch.Tensor]:
    """utils.get_bce_weight (utils.py:726-749), vectorised; None stands for the all-ones weight."""
    if not balance_positive_negative:
        return None
    pos = output_label_pts.bool()
    pp = pos.float().mean(dim=2, keepdim=True)
    w = torch.where(pos, 1.0 / (pp + 1e-10), 1.0 / ((1 - pp) + 1e-10))
    return w * (float(w.numel()) / w.sum())


def get_losses_ovssc(net, batch: dict, **kwargs):
    """= semabs_b200.train_ovssc.get_losses (the reference's train_ovssc.get_losses contract: stats + per-cutoff DataFrame)."""
    from .train_ovssc import get_losses

    batch.setdefault("scene_id", [f"scene_{i}" for i in range(batch["output_label_pts"].shape[0])])
    return get_losses(net, batch, **kwargs)


def get_losses_vool(net, batch: dict, **kwargs):
    """= semabs_b200.train_vool.get_losses."""
    from .train_vool import get_losses

    B, D = batch["output_label_pts"].shape[:2]
    batch.setdefault("scene_id", [f"scene_{i}" for i in range(B)])
    for k in ("target_obj_name", "reference_obj_name"):
        batch.setdefault(k, [[""] * B for _ in range(D)])
    return get_losses(net, batch, **kwargs)


def train_step(net, batch: dict, get_losses_fn, optimizer, grad_max_norm: float = 1e5, lr_scheduler=None, **kwargs):
    """One iteration of utils.loop's train branch (utils.py:404-422): losses -> zero_grad -> backward ->
    [DDP gradient averaging] -> clip_grad_norm_ -> optimizer.step -> lr_scheduler.step -> steps += 1 -> `gradnorm` stat.
    With this module's Lamb the clip is folded into the optimiser sweep (the gradients themselves stay unclipped in
    memory; `gradnorm` is reported as what utils.compute_grad_norm would see AFTER the reference's in-place clip)."""
    stats, _ = get_losses_fn(net=net, batch=batch, **kwargs)
    optimizer.zero_grad(set_to_none=True)
    with GradientBuckets() as buckets:
        stats["loss"].backward()
    all_reduce_gradients(net.parameters(), skip=buckets.reduced)
    if isinstance(optimizer, Lamb):
        optimizer.step(max_grad_norm=grad_max_norm)
        total = optimizer.last_grad_norm
        stats["gradnorm"] = total * torch.clamp(grad_max_norm / (total + 1e-6), max=1.0)
    else:
        total = clip_grad_norm_(net.parameters(), grad_max_norm)
        optimizer.step()
        bump_weights_epoch()
        stats["gradnorm"] = total * torch.clamp(grad_max_norm / (total + 1e-6), max=1.0)
    if lr_scheduler is not None:
        lr_scheduler.step()
    net.steps += 1
    return stats


class _ChunkTable:
    """Device table of (p, g, m, v, n, tensor) records for a fixed list of parameter tensors."""

    def __init__(self, params: List[torch.Tensor], grads: List[torch.Tensor], ms: List[torch.Tensor], vs: List[torch.Tensor]):
        assert lib().semabs_lamb_chunk_bytes() == 40
        rec = []
        for t, (p, g, m, v) in enumerate(zip(params, grads, ms, vs)):
            assert p.is_contiguous() and g.is_contiguous() and p.dtype == torch.float32 and g.dtype == torch.float32
            n = p.numel()
            for off in range(0, n, CHUNK):
                k = min(CHUNK, n - off)
                rec.append(struct.pack("<QQQQii", p.data_ptr() + 4 * off, g.data_ptr() + 4 * off,
                                       m.data_ptr() + 4 * off if m is not None else 0,
                                       v.data_ptr() + 4 * off if v is not None else 0, k, t))
        self.n_chunks = len(rec)
        self.n_tensors = len(params)
        dev = params[0].device
        self.table = torch.frombuffer(bytearray(b"".join(rec)), dtype=torch.uint8).to(dev)
        self.key = _table_key(params, grads, ms, vs)


def _table_key(params, grads, ms, vs):
    """identity of everything a table points at: parameter, gradient AND moment buffers (optimizer.load_state_dict replaces
    the moments without touching the parameters)"""
    dp = lambda t: 0 if t is None else t.data_ptr()
    return tuple((dp(p), dp(g), dp(m), dp(v)) for p, g, m, v in zip(params, grads, ms, vs))


def grad_sumsq(table: _ChunkTable) -> torch.Tensor:
    out = torch.empty(1, dtype=torch.float64, device=table.table.device)
    check(lib().semabs_grad_sumsq(ptr(table.table), i32(table.n_chunks), ptr(out), stream_ptr()))
    return out


def clip_grad_norm_(parameters: Iterable[torch.Tensor], max_norm: float) -> torch.Tensor:
    """Drop-in for torch.nn.utils.clip_grad_norm_(params, max_norm) (L2): returns the total norm (device tensor)."""
    ps = [p for p in parameters if p.grad is not None]
    table = _ChunkTable([p.data for p in ps], [p.grad.data for p in ps], [None] * len(ps), [None] * len(ps))
    ss = grad_sumsq(table)
    check(lib().semabs_clip_grads(ptr(table.table), i32(table.n_chunks), ptr(ss), f32(max_norm), stream_ptr()))
    return ss.sqrt().float()[0]


class Lamb(torch.optim.Optimizer):
    """Same constructor and state layout (`step`, `exp_avg`, `exp_avg_sq`) as the reference Lamb; `step()` runs one
    fused multi-tensor update. Pass `max_grad_norm` to fold clip_grad_norm_ into the same sweep."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, adam=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        self.adam = adam
        self._tables = {}
        self.last_grad_norm = None
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}  # the moment buffers were replaced: never keep pointers into the old ones

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        loss = closure() if closure is not None else None
        work = []
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]  # params without grad are skipped (lamb.py:71-72)
            if not ps:
                continue
            for p in ps:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p.data)
                    st["exp_avg_sq"] = torch.zeros_like(p.data)
                st["step"] += 1
            params, grads = [p.data for p in ps], [p.grad.data for p in ps]
            ms, vs = [self.state[p]["exp_avg"] for p in ps], [self.state[p]["exp_avg_sq"] for p in ps]
            tb = self._tables.get(gi)
            if tb is None or tb.key != _table_key(params, grads, ms, vs):
                tb = self._tables[gi] = _ChunkTable(params, grads, ms, vs)
            work.append((group, tb))
        ss = None
        if max_grad_norm is not None and work:
            # clip_grad_norm_(net.parameters(), max) uses the GLOBAL norm over every parameter with a gradient
            # (utils.py:415): one sum of squares over all groups, the same scalar for every group's sweep
            ss = grad_sumsq(work[0][1])
            for _, tb in work[1:]:
                ss = ss + grad_sumsq(tb)
            self.last_grad_norm = ss.sqrt().float()[0]
        for group, tb in work:
            norms = torch.empty(2 * tb.n_tensors, dtype=torch.float64, device=tb.table.device)
            b1, b2 = group["betas"]
            check(lib().semabs_lamb_step(ptr(tb.table), i32(tb.n_chunks), i32(tb.n_tensors), ptr(norms), ptr(ss),
                                         f32(max_grad_norm or 0.0), f32(group["lr"]), f32(b1), f32(b2), f32(group["eps"]),
                                         f32(group["weight_decay"]), i32(int(self.adam)), stream_ptr()))
            self.last_norms = norms
        # the update went through raw device pointers: torch's version counters did not move, so tell the modules that
        # cache fp16 MMA-operand packs of their weights (ResidualUNet3D._packed, UNetBackward._packed) to rebuild them
        bump_weights_epoch()
        return loss


class GradientBuckets:
    """Per-level gradient buckets, all-reduced while the backward pass is still running (the overlap
    DistributedDataParallel's reducer gives the reference at utils.py:255-258).  `with GradientBuckets():` around
    `loss.backward()` makes the UNet's backward (unet3d_bwd._UNetFn) hand over the weight gradients of each resolution
    level as soon as that level's weight-gradient kernels are queued: one flat NCCL all-reduce per level, asynchronous
    (NCCL's stream waits for the producing kernels, the compute stream carries on with the next level), joined once at
    the end of the node's backward.  The node then returns the AVERAGED gradients, and `reduced` names the parameters
    `all_reduce_gradients` must leave alone.  Inactive (a no-op context) without an initialised process group, with one
    rank, or with SEMABS_GRAD_BUCKETS=0.  Every rank runs the same levels in the same order, so the collectives match."""

    _active: Optional["GradientBuckets"] = None

    def __init__(self, world_size: Optional[int] = None, keep_local: bool = False):
        self.local: Optional[dict] = {} if keep_local else None  # debugging: this rank's own gradients, before the reduce
        on = dist.is_available() and dist.is_initialized() and os.environ.get("SEMABS_GRAD_BUCKETS", "1") != "0"
        self.world = (world_size or dist.get_world_size()) if on else 1
        self.enabled = on and self.world > 1
        self.pending: list = []
        self.reduced: set = set()
        self.n_buckets = 0

    def __enter__(self):
        if self.enabled:
            GradientBuckets._active = self
        return self

    def __exit__(self, *exc):
        GradientBuckets._active = None
        assert not self.pending or exc[0] is not None, "GradientBuckets: a bucket was submitted but never joined"
        return False

    @staticmethod
    def active() -> Optional["GradientBuckets"]:
        return GradientBuckets._active

    def submit(self, grads: dict, keys: list):
        """Start the all-reduce of grads[k] for k in keys (one flat bucket)."""
        if not keys:
            return
        flat = torch.cat([grads[k].reshape(-1) for k in keys])
        if self.local is not None:
            self.local.update((k, grads[k].clone()) for k in keys)
        work = dist.all_reduce(flat, async_op=True)
        self.pending.append((work, flat, keys))
        self.n_buckets += 1

    def join(self, grads: dict):
        """Wait for every bucket (stream-side on NCCL) and replace grads[k] by views of the averaged buckets."""
        for work, flat, keys in self.pending:
            work.wait()
            flat.mul_(1.0 / self.world)
            off = 0
            for k in keys:
                n = grads[k].numel()
                grads[k] = flat[off : off + n].view_as(grads[k])
                off += n
                self.reduced.add(k)
        self.pending.clear()


def all_reduce_gradients(parameters: Iterable[torch.Tensor], world_size: Optional[int] = None, skip=()):
    """Data-parallel gradient averaging over NCCL (what DistributedDataParallel(find_unused_parameters=True) does at
    utils.py:255-258): one flat bucket per call.  A parameter that has a gradient on ANY rank ends up with the averaged
    gradient on EVERY rank (ranks without one contribute zeros); a parameter unused everywhere keeps grad None on all
    ranks — so LAMB's skip decision (arm/optim/lamb.py:71-72) is identical across ranks and the replicas stay in sync
    (SURVEY.md §8e).  The reference sets NCCL_P2P_DISABLE=1 (utils.py:132); we do not.  The UNet's parameters normally
    arrive here already averaged (GradientBuckets, overlapped with the backward pass) and are passed in `skip`."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = world_size or dist.get_world_size()
    done = {id(p) for p in skip}  # already averaged by GradientBuckets during backward
    ps = [p for p in parameters if id(p) not in done]
    if not ps:
        return
    dev = ps[0].device
    has = torch.tensor([0.0 if p.grad is None else 1.0 for p in ps], device=dev)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps] + [has])
    dist.all_reduce(flat)
    used = (flat[-len(ps):] > 0).tolist()
    flat /= world
    off = 0
    for p, u in zip(ps, used):
        n = p.numel()
        if u:
            g = flat[off : off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
        off += n
