"""Ensemble of checkpoints — the reference's ``src/tasks/gqa_ensemble.py`` flow on the device.

The reference runs one full sweep per checkpoint, keeps the per-class sigmoid vector of every question on the host
(``ood_evaluate`` :95-108), combines them question by question in a Python loop (``mean`` / ``multiply`` :165-171) and
takes ``np.argmax`` (:173-177).  Here every checkpoint's logits of a batch stay on the device and ONE kernel
(``rgqa_score_ensemble``) does sigmoid + combine + arg-max; what leaves the GPU per question is (score, label), as for the
single-model scorers.  Same state-dict loading convention as the reference (``load`` :127-133: ``.module`` stripped).
"""
from __future__ import annotations

import torch

from .. import ops
from .gqa_data import GQAOODEvaluator
from .gqa_eval import build_quesid2ans, gather_results


def strip_module_keys(state_dict):
    """gqa_ensemble.py:129-131 — checkpoints saved under nn.DataParallel carry '.module' in their keys."""
    return {k.replace(".module", ""): v for k, v in state_dict.items()}


class GQAEnsemble:
    """``models``: GQAModel instances (one per checkpoint of ``--load a,b,...``), all on the same device."""

    def __init__(self, models, dataset, ensemble_method="mean", tau=0.5):
        if ensemble_method not in ("mean", "multiply"):
            raise ValueError(f"Ensemble method {ensemble_method} not implemented!")  # gqa_ensemble.py:171
        if not 1 <= len(models) <= 8:
            raise ValueError("between 1 and 8 checkpoints")
        self.models, self.dataset, self.method, self.tau = list(models), dataset, ensemble_method, tau

    def score_device_batch(self, feats, boxes, ids, mask, seg=None):
        """(score, label) of the ensemble for one device batch: K forwards, one combine kernel."""
        logits = [m.forward_ids(feats, boxes, ids, seg, mask)[0] for m in self.models]
        return ops.score_ensemble(logits, self.method)

    def ood_evaluate(self, loader, dump=None, gather=True):
        """loader yields (ques_ids, feats, boxes, ids, mask) host batches of this rank's shard, in order."""
        dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        qids, scores, labels = [], [], []
        for ques_id, feats, boxes, ids, mask in loader:
            f, b, i, m = (t.to(dev, non_blocking=True) for t in (feats, boxes, ids, mask))
            s, l = self.score_device_batch(f, b, i, m)
            qids.extend(ques_id)
            scores.append(s)
            labels.append(l)
        score = torch.cat(scores) if scores else torch.empty(0, device=dev)
        label = torch.cat(labels) if labels else torch.empty(0, dtype=torch.int32, device=dev)
        score, label, qids = gather_results(score, label, qids) if gather else (score.cpu(), label.cpu(), qids)
        if score is None:
            return None
        quesid2ans = build_quesid2ans(qids, label.numpy(), score.numpy(), self.dataset.label2ans)
        evaluator = GQAOODEvaluator(self.dataset, self.tau)
        results = evaluator.evaluate(quesid2ans)
        if dump is not None:
            evaluator.dump_result(quesid2ans, dump)
        return results
