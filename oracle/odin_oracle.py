"""ORACLE — test infrastructure, not product code.

CPU restatement (torch autograd through the oracle's own LXMERT forward) of the two scorers that need input gradients:
ODIN (``src/tasks/gqa_odin.py:97-131``) and the Mahalanobis score with input pre-processing
(``src/tasks/gqa_maha.py:216-276``).  The lines are restated in order, with the reference's variable names.

**Pinned**: ``tests/test_oracle_cpu.py::test_scorer_oracles_pinned_to_reference_drivers`` runs both functions with the tiny
model of ``tests/golden/tiny_vqa.py`` (``model=``) and compares with what the UNMODIFIED reference drivers
(``gqa_odin.GQA.ood_evaluate``, ``gqa_maha.GQA.ood_evaluate`` with ``--noise`` > 0) returned for the same model and inputs
(``tests/golden/make_golden_scorers.py`` -> ``scorers_reference.npz``).  The LXMERT forward they differentiate by default is
pinned separately (``lxmert_oracle``)."""
from __future__ import annotations

import torch
from torch import nn

from . import lxmert_oracle as O


def _model(sd, feats, boxes, ids, seg, mask):
    x = O.lxmert_encode(sd, feats, boxes, ids, seg, mask)
    return O.gqa_head(sd, x), x


def _forward(sd, ids, seg, mask, model):
    """(feats, boxes) -> (logit, x): the oracle LXMERT on ``sd`` by default, or any callable standing in for ``self.model``."""
    if model is not None:
        return model
    return lambda feats, boxes: _model(sd, feats, boxes, ids, seg, mask)


def odin(sd, feat, pos, ids, seg, mask, temperature=1.0, noise=1e-4, model=None):
    """gqa_odin.py:97-131.  Returns (score, label, feats.grad, boxes.grad)."""
    fwd = _forward(sd, ids, seg, mask, model)
    bceloss = nn.BCEWithLogitsLoss(reduction="none")
    feats, boxes = feat.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    logit, _ = fwd(feats, boxes)
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
        logit, _ = fwd(featsPt, boxesPt)
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


def mahalanobis_with_noise(sd, feat, pos, ids, seg, mask, sample_mean, precision, noise=1e-4, model=None):
    """gqa_maha.py:216-276 with ``--noise`` > 0.  Returns (score, label, feats.grad, boxes.grad)."""
    fwd = _forward(sd, ids, seg, mask, model)
    feats, boxes = feat.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    _, outputs = fwd(feats, boxes)
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
        _, noised_outputs = fwd(featsPt, boxesPt)
        score, label = torch.max(_gaussian_scores(noised_outputs.data, sample_mean, precision), dim=1)
    return score, label, feats.grad.detach(), boxes.grad.detach()
