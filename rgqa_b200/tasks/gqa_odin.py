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
