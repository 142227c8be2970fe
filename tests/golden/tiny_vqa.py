"""A tiny differentiable stand-in for ``GQAModel`` / ``GQAModel_maha`` with the reference's call signature
``model(feats, boxes, sent) -> logit`` (or ``(logit, x)``).

It exists so that the reference's scorer DRIVERS (``src/tasks/gqa_energy.py``, ``gqa_maha.py``, ``gqa_odin.py``,
``gqa_conf.py``, ``gqa_ensemble.py`` — the loops around the model call, which are the lines SURVEY §8 a15-a17 / (f) name) can be
executed unmodified on a few dozen synthetic questions in seconds, with and without input gradients, and their outputs
committed as golden vectors (``make_golden_scorers.py``).  The LXMERT forward itself is pinned separately
(``make_golden.py``); here only the code around it is under test, so the model is as small as possible.
Weights live in the golden file, not in an RNG."""
import torch
from torch import nn


class TinyVQA(nn.Module):
    def __init__(self, feat_dim=32, hid=16, num_classes=12, maha=False):
        super().__init__()
        self.fc_feat = nn.Linear(feat_dim, hid)
        self.fc_box = nn.Linear(4, hid)
        self.len_emb = nn.Embedding(8, hid)
        self.head = nn.Linear(hid, num_classes)
        self.maha = maha

    def forward(self, feats, boxes, sent):
        lens = torch.tensor([len(s.split()) % 8 for s in sent], dtype=torch.long, device=feats.device)
        x = torch.tanh(self.fc_feat(feats).mean(1) + self.fc_box(boxes).mean(1) + self.len_emb(lens))
        logit = self.head(x) * 4.0
        return (logit, x) if self.maha else logit
