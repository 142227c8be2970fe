"""ORACLE — test infrastructure, not product code.

CPU restatement (torch autograd through the oracle's own LXMERT forward) of the two scorers that need input gradients:
ODIN (``src/tasks/gqa_odin.py:97-131``) and the Mahalanobis score with input pre-processing
(``src/tasks/gqa_maha.py:216-276``).  The drivers cannot be imported (``vilt`` -> ``timm``), so the lines are restated in
order, with the reference's variable names.  **Parity unpinned** beyond the pinned forward they differentiate (the
reference ships no test for these flows)."""
from __future__ import annotations

import torch
from torch import nn

from . import lxmert_oracle as O


def _model(sd, feats, boxes, ids, seg, mask):
    x = O.lxmert_encode(sd, feats, boxes, ids, seg, mask)
    return O.gqa_head(sd, x), x


def odin(sd, feat, pos, ids, seg, mask, temperature=1.0, noise=1e-4):
    """gqa_odin.py:97-131.  Returns (score, label, feats.grad, boxes.grad)."""
    bceloss = nn.BCEWithLogitsLoss(reduction="none")
    feats, boxes = feat.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    logit, _ = _model(sd, feats, boxes, ids, seg, mask)
    preds = torch.sigmoid(logit / temperature)                       # :104
    labels = (torch.ones(preds.shape) * (preds >= 0.5)).float()      # :106-107
    loss = bceloss(logit, labels)                                    # :110
    idx = torch.max(preds, dim=1)[1].unsqueeze(-1)                   # :112
    loss = torch.mean(torch.gather(loss, 1, idx))                    # :113
    loss.backward()
    gradient = (torch.ge(feats.grad.data, 0).float() - 0.5) * 2      # :118-119
    featsPt = torch.add(feats.data, gradient, alpha=-noise)
    gradient = (torch.ge(boxes.grad.data, 0).float() - 0.5) * 2
    boxesPt = torch.add(boxes.data, gradient, alpha=-noise)
    with torch.no_grad():
        logit, _ = _model(sd, featsPt, boxesPt, ids, seg, mask)
        outputs = torch.sigmoid(logit / temperature)
        score, label = outputs.max(1)
    return score, label, feats.grad.detach(), boxes.grad.detach()


def _gaussian_scores(x, sample_mean, precision):
    out = 0
    for i in range(sample_mean.shape[0]):                            # gqa_maha.py:228-236 / :258-266
        zero_f = x - sample_mean[i]
        term_gau = -0.5 * torch.mm(torch.mm(zero_f, precision), zero_f.t()).diag()
        out = term_gau.view(-1, 1) if i == 0 else torch.cat((out, term_gau.view(-1, 1)), 1)
    return out


def mahalanobis_with_noise(sd, feat, pos, ids, seg, mask, sample_mean, precision, noise=1e-4):
    """gqa_maha.py:216-276 with ``--noise`` > 0.  Returns (score, label, feats.grad, boxes.grad)."""
    feats, boxes = feat.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    _, outputs = _model(sd, feats, boxes, ids, seg, mask)
    gaussian_score = _gaussian_scores(outputs.data, sample_mean, precision)
    sample_pred = gaussian_score.max(1)[1]                           # :239
    batch_sample_mean = sample_mean.index_select(0, sample_pred)
    zero_f = outputs - batch_sample_mean
    pure_gau = -0.5 * torch.mm(torch.mm(zero_f, precision), zero_f.t()).diag()
    loss = torch.mean(-pure_gau)
    loss.backward()
    gradient = (torch.ge(feats.grad.data, 0).float() - 0.5) * 2
    featsPt = torch.add(feats.data, gradient, alpha=-noise)
    gradient = (torch.ge(boxes.grad.data, 0).float() - 0.5) * 2
    boxesPt = torch.add(boxes.data, gradient, alpha=-noise)
    with torch.no_grad():
        _, noised_outputs = _model(sd, featsPt, boxesPt, ids, seg, mask)
        score, label = torch.max(_gaussian_scores(noised_outputs.data, sample_mean, precision), dim=1)
    return score, label, feats.grad.detach(), boxes.grad.detach()
