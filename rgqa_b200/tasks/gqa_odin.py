"""The two scorers that perturb the input along the sign of a gradient — ODIN (``src/tasks/gqa_odin.py:97-131``) and
the Mahalanobis score with ``--noise`` > 0 (``src/tasks/gqa_maha.py:216-276``) — on the GPU.

Both are: forward, a scalar loss on the model's own prediction, ``backward()`` down to ``feats`` / ``boxes``, one step of
size ``noise`` against the gradient's sign, a second forward, the score.  ``LXMERTInputGrad`` (fp32 path) supplies the
taped forward and d loss / d (feats, boxes); everything per-element runs in this package's kernels, torch only holds
tensors and does the B-element bookkeeping (which logit, which class mean)."""
from __future__ import annotations

import torch

from .. import _lib as L
from .. import ops


def _step_against_gradient_sign(x, grad, noise):
    """x - noise * sign, sign = +1 where grad >= 0 else -1 (gqa_odin.py:118-120): one fused elementwise pass."""
    x, grad = x.contiguous(), grad.contiguous()
    y = torch.empty_like(x)
    L.check(L.load().rgqa_sign_step(L.current_stream_ptr(), x.data_ptr(), grad.data_ptr(), y.data_ptr(), x.numel(), float(noise)),
            "rgqa_sign_step")
    return y


def odin_scores(ig, feat, pos, ids, seg, mask, temperature=1.0, noise=1e-4):
    """(score f32 [B], label i32 [B]) — max sigmoid(logit / T) after the ODIN input pre-processing."""
    feat, pos = feat.to(ig.dev, torch.float32), pos.to(ig.dev, torch.float32)
    logits, _ = ig.forward(feat, pos, ids, seg, mask)
    B = logits.shape[0]
    _, idx = ops.score_logits(logits, "msp")                            # arg-max of sigmoid(logit / T) = arg-max of sigmoid(logit)
    z = logits.gather(1, idx.long().unsqueeze(1)).squeeze(1)            # the B logits the loss looks at (:112-113)
    label = (torch.sigmoid(z / temperature) >= 0.5).float()             # :106
    dlogits = torch.zeros_like(logits)
    dlogits.scatter_(1, idx.long().unsqueeze(1), ((torch.sigmoid(z) - label) / B).unsqueeze(1))   # d mean(BCE) / d logit
    dfeat, dpos = ig.backward(dlogits=dlogits)
    logits2, _ = ig.forward(_step_against_gradient_sign(feat, dfeat, noise), _step_against_gradient_sign(pos, dpos, noise), ids, seg, mask)
    score, label2 = ops.score_logits(logits2 / temperature, "msp")
    return score, label2, dfeat, dpos


def mahalanobis_noise_scores(ig, feat, pos, ids, seg, mask, sample_mean, precision, noise=1e-4):
    """(score, label) of gqa_maha.py:216-276 with noise > 0: loss = mean_b 1/2 z^T P z, z = x - mean of the predicted class."""
    feat, pos = feat.to(ig.dev, torch.float32), pos.to(ig.dev, torch.float32)
    mu = sample_mean.to(ig.dev, torch.float32).contiguous()
    Pm = precision.to(ig.dev, torch.float32).contiguous()
    _, x = ig.forward(feat, pos, ids, seg, mask)
    B = x.shape[0]
    _, pred = ops.mahalanobis(x, mu, Pm)
    zf = (x - mu.index_select(0, pred.long())).contiguous()
    # d/dx of 1/2 z^T P z = 1/2 (P + P^T) z ; the mean over the batch divides by B
    psym = (0.5 * (Pm + Pm.t())).contiguous()
    dpooled = ops.linear_f32(zf, psym) / B
    dfeat, dpos = ig.backward(dpooled=dpooled)
    _, x2 = ig.forward(_step_against_gradient_sign(feat, dfeat, noise), _step_against_gradient_sign(pos, dpos, noise), ids, seg, mask)
    score, label = ops.mahalanobis(x2, mu, Pm)
    return score, label, dfeat, dpos


class GraphedInputGradScorer:
    """ODIN / Mahalanobis-with-noise for a FIXED batch shape as ONE CUDA-graph replay.

    A scoring pass is forward + input-gradient backward + sign step + forward: ~1 200 kernel launches issued from Python, which
    at the reference scripts' batch size (32) costs several times the GPU work.  The first call of a shape runs eagerly (lazy
    kernel attributes, weight copies), the second captures the whole pass over static input buffers, later calls copy the batch
    in and replay.  The returned (score, label) are the graph's own output tensors: consume (or clone) them before the next call.
    Same arithmetic as :func:`odin_scores` / :func:`mahalanobis_noise_scores`, kernel for kernel."""

    def __init__(self, ig, kind="odin", temperature=1.0, noise=1e-4, sample_mean=None, precision=None):
        if kind not in ("odin", "maha"):
            raise ValueError(f"unknown scorer {kind!r} (odin | maha)")
        if kind == "maha" and (sample_mean is None or precision is None):
            raise ValueError("the Mahalanobis scorer needs sample_mean and precision")
        self.ig, self.kind, self.temperature, self.noise = ig, kind, temperature, noise
        # device-resident statistics: nothing may cross the PCIe bus under capture
        self.mu = None if sample_mean is None else sample_mean.to(ig.dev, torch.float32).contiguous()
        self.P = None if precision is None else precision.to(ig.dev, torch.float32).contiguous()
        self._graphs = {}
        self.captures = 0

    def _run(self, feat, pos, ids, seg, mask):
        if self.kind == "odin":
            s, l, _, _ = odin_scores(self.ig, feat, pos, ids, seg, mask, self.temperature, self.noise)
        else:
            s, l, _, _ = mahalanobis_noise_scores(self.ig, feat, pos, ids, seg, mask, self.mu, self.P, self.noise)
        return s, l

    def __call__(self, feat, pos, ids, seg, mask):
        dev = self.ig.dev
        args = [feat.to(dev, torch.float32), pos.to(dev, torch.float32), ids.to(dev), seg.to(dev), mask.to(dev)]
        key = tuple((tuple(t.shape), t.dtype) for t in args)
        st = self._graphs.get(key)
        if st is None:                                   # first sight of a shape: eager
            self._graphs[key] = {}
            return self._run(*args)
        if "graph" not in st:
            st["inputs"] = [torch.empty_like(t) for t in args]
            for dst, src in zip(st["inputs"], args):
                dst.copy_(src)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st["out"] = self._run(*st["inputs"])
            st["graph"] = g
            self.captures += 1
        for dst, src in zip(st["inputs"], args):
            dst.copy_(src, non_blocking=True)
        st["graph"].replay()
        return st["out"]
