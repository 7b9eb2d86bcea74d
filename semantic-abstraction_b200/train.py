"""Optimiser side of the reference training step on the B200 kernels: masked BCE-with-logits (loss + gradient),
gradient-norm clipping and LAMB — mirrors of `train_ovssc.get_losses`' loss (train_ovssc.py:133-150),
`torch.nn.utils.clip_grad_norm_` as used at utils.py:415 and `arm.optim.lamb.Lamb` (arm/optim/lamb.py:25-127),
plus the data-parallel gradient all-reduce (the one real exchange step of the path, utils.py:255-258).

NOT here yet: the backward kernels of SemAbs3D / ResidualUNet3D that produce the gradients.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import check, f32, i32, lib, ptr, stream_ptr

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
        self.key = tuple((p.data_ptr(), g.data_ptr()) for p, g in zip(params, grads))


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
        self._table = None
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
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
            key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
            if self._table is None or self._table.key != key:
                self._table = _ChunkTable([p.data for p in ps], [p.grad.data for p in ps],
                                          [self.state[p]["exp_avg"] for p in ps], [self.state[p]["exp_avg_sq"] for p in ps])
            tb = self._table
            norms = torch.empty(2 * tb.n_tensors, dtype=torch.float64, device=tb.table.device)
            ss = grad_sumsq(tb) if max_grad_norm is not None else None
            b1, b2 = group["betas"]
            check(lib().semabs_lamb_step(ptr(tb.table), i32(tb.n_chunks), i32(tb.n_tensors), ptr(norms), ptr(ss),
                                         f32(max_grad_norm or 0.0), f32(group["lr"]), f32(b1), f32(b2), f32(group["eps"]),
                                         f32(group["weight_decay"]), i32(int(self.adam)), stream_ptr()))
            self.last_norms = norms
        return loss


def all_reduce_gradients(parameters: Iterable[torch.Tensor], world_size: Optional[int] = None):
    """Data-parallel gradient averaging over NCCL (what DistributedDataParallel does at utils.py:255-258): one flat
    bucket per call; parameters without a gradient contribute zeros on every rank so that all ranks make the same
    skip decision. The reference sets NCCL_P2P_DISABLE=1 (utils.py:132); we do not."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = world_size or dist.get_world_size()
    ps = list(parameters)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
    dist.all_reduce(flat)
    flat /= world
    off = 0
    for p in ps:
        n = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off : off + n].view_as(p))
        off += n
