"""BertAdam on the device — ``src/lxrt/optimization.py`` of the reference (schedules :24-49, ``BertAdam`` :52-180) with the
update in one CUDA kernel per tensor (``rgqa_bertadam_step``) and the reference trainer's gradient clipping
(``nn.utils.clip_grad_norm_(model.parameters(), 5.)``, ``src/tasks/gqa_conf.py:197``) folded into it.

Groundwork for SURVEY §8 a22 / config 5: the gradients still have to come from somewhere — the backward graph of the
encoder is not built yet, so today this optimiser serves tensors whose ``.grad`` the caller provides.
"""
from __future__ import annotations

import math

import torch

from .. import _lib as L


def _ramp(x, warmup):
    """Linear ramp 0 -> 1 over the first `warmup` fraction of training (shared by the three schedules)."""
    return x / warmup


def warmup_cosine(x, warmup=0.002):
    """optimization.py:24-28: ramp, then half a cosine period down to 0 at x = 1."""
    return _ramp(x, warmup) if x <= warmup else 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x, warmup=0.002):
    """optimization.py:30-35: ramp, then 1."""
    return _ramp(x, warmup) if x < warmup else 1.0


def warmup_linear(x, warmup=0.002):
    """optimization.py:37-42: ramp, then a straight line that reaches 0 at x = 1 (and stays there)."""
    return _ramp(x, warmup) if x < warmup else max((x - 1.0) / (warmup - 1.0), 0)


SCHEDULES = {"warmup_cosine": warmup_cosine, "warmup_constant": warmup_constant, "warmup_linear": warmup_linear}


class BertAdam:
    """Same constructor arguments and state as the reference optimiser (``next_m``, ``next_v``, per-tensor ``step``;
    the learning rate of update k uses ``step = k`` BEFORE the increment, :158-160, :171).  ``params``: CUDA fp32 tensors
    updated in place.  ``step(grads, clip_norm)``: ``grads`` parallel to ``params`` (None entries are skipped like
    ``p.grad is None``); ``clip_norm > 0`` applies the global-norm clip of the reference trainer before the update."""

    def __init__(self, params, lr, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 weight_decay=0.01, max_grad_norm=1.0):
        checks = (("lr", lr >= 0.0), ("schedule", schedule in SCHEDULES), ("warmup", warmup == -1 or 0.0 <= warmup < 1.0),
                  ("b1", 0.0 <= b1 < 1.0), ("b2", 0.0 <= b2 < 1.0), ("e", e >= 0.0))
        for name, ok in checks:           # same admissible ranges as the reference constructor (optimization.py:66-79)
            if not ok:
                raise ValueError(f"BertAdam: argument {name} out of range")
        self.params = list(params)
        for p in self.params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError("BertAdam: contiguous CUDA fp32 tensors required (there is no CPU path)")
        self.group = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                          weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        self.state = [dict(step=0, next_m=torch.zeros_like(p), next_v=torch.zeros_like(p)) for p in self.params]

    def get_lr(self):
        g = self.group
        out = []
        for st in self.state:
            if g["t_total"] != -1:
                out.append(g["lr"] * SCHEDULES[g["schedule"]](st["step"] / g["t_total"], g["warmup"]))
            else:
                out.append(g["lr"])
        return out

    def step(self, grads, clip_norm=0.0):
        lib = L.load()
        g = self.group
        stream = L.current_stream_ptr()
        live = [(p, gr, st) for p, gr, st in zip(self.params, grads, self.state) if gr is not None]
        sumsq = None
        if clip_norm > 0 and live:
            sumsq = torch.zeros(1, dtype=torch.float64, device=live[0][0].device)
            for p, gr, _ in live:
                L.check(lib.rgqa_grad_sumsq(stream, L.ptr(gr), gr.numel(), L.ptr(sumsq)), "rgqa_grad_sumsq")
        for p, gr, st in live:
            assert gr.is_cuda and gr.dtype == torch.float32 and gr.is_contiguous() and gr.shape == p.shape
            if g["t_total"] != -1:
                lr = g["lr"] * SCHEDULES[g["schedule"]](st["step"] / g["t_total"], g["warmup"])
            else:
                lr = g["lr"]
            L.check(lib.rgqa_bertadam_step(stream, L.ptr(p), L.ptr(gr), L.ptr(st["next_m"]), L.ptr(st["next_v"]), p.numel(),
                                           float(lr), g["b1"], g["b2"], g["e"], g["weight_decay"], L.ptr(sumsq),
                                           float(clip_norm)), "rgqa_bertadam_step")
            st["step"] += 1
        return None if sumsq is None else sumsq


def bce_loss_and_grad(logit, target):
    """(loss, dlogit): ``BCEWithLogitsLoss()(logit, target) * logit.size(1)`` of the reference trainer
    (src/tasks/gqa_conf.py:193-194) and its gradient w.r.t. the logits, in one kernel."""
    assert logit.is_cuda and logit.dtype == target.dtype == torch.float32 and logit.shape == target.shape
    logit, target = logit.contiguous(), target.contiguous()
    B, A = logit.shape
    dlogit = torch.empty_like(logit)
    loss = torch.zeros(1, dtype=torch.float64, device=logit.device)
    lib = L.load()
    L.check(lib.rgqa_bce_loss_grad(L.current_stream_ptr(), L.ptr(logit), L.ptr(target), B, A, L.ptr(dlogit), L.ptr(loss)),
            "rgqa_bce_loss_grad")
    return loss, dlogit
