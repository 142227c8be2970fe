"""Host logic of the evaluation driver mirror (no GPU): ``GQA.ood_evaluate`` called with the REFERENCE's signature —
``ood_evaluate(eval_tuple, dump=)`` / ``ood_evaluate(eval_tuple, (mean, precision), dump=)`` — on the loader batches the unmodified
reference drivers were run on (tests/golden/make_golden_scorers.py), with the tiny stand-in model in place of the CUDA model and
the oracle's scorers in place of the device scorers (scaffolding: the loop, the evaluator and the wire file are what is under
test).  The result dicts must be the reference's, float for float."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class _FakeEngine:
    def __init__(self):
        self.maha = None

    def set_maha(self, mean, precision):
        self.maha = (mean, precision)

    def score(self, method, logits=None, pooled=None):
        from oracle import scorers_oracle as S
        if method == "maha":
            return S.mahalanobis(pooled, *self.maha)[:2]
        return {"msp": S.msp, "energy": S.energy_top2}[method](logits)


class _FakeModel:
    def __init__(self, net):
        self.net, self._engine = net, _FakeEngine()

    def parameters(self):
        return self.net.parameters()

    def engine(self):
        return self._engine

    def __call__(self, feats, boxes, sent):
        with torch.no_grad():
            return self.net(feats, boxes, sent)


@pytest.mark.parametrize("method", ["msp", "energy", "maha"])
def test_reference_signature_ood_evaluate_reproduces_reference_results(tmp_path, method):
    sys.path.insert(0, GOLD)
    from tiny_vqa import TinyVQA
    from rgqa_b200.tasks.gqa_conf import GQA
    z = np.load(os.path.join(GOLD, "scorers_reference.npz"))
    with open(os.path.join(GOLD, "scorers_reference.json")) as f:
        meta = json.load(f)
    sd = {k[len("model."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("model.")}
    C, D = sd["head.weight"].shape
    net = TinyVQA(sd["fc_feat.weight"].shape[1], D, C, maha=True)
    net.load_state_dict(sd, strict=True)
    gqa = object.__new__(GQA)                       # the constructor builds the CUDA model; only the driver loop is under test
    gqa.args = types.SimpleNamespace(method=method, tau=0.5)
    gqa.model = _FakeModel(net.eval())
    qids, sents = meta["qids_test"], meta["sents_test"]
    feats, boxes, target = (torch.from_numpy(z[k]) for k in ("feat_test", "box_test", "target_test"))
    loader = [(qids[i:i + 16], feats[i:i + 16], boxes[i:i + 16], sents[i:i + 16], target[i:i + 16]) for i in range(0, len(qids), 16)]
    dset = types.SimpleNamespace(label2ans=meta["label2ans"], num_answers=len(meta["label2ans"]),
                                 id2datum={q: {"label": {a: v for a, v in pairs}} for q, pairs in zip(qids, meta["labels_test"])})
    dump = str(tmp_path / f"{method}_predict.json")
    if method == "maha":       # gqa_maha.py:191: ood_evaluate(eval_tuple, pack, dump=)
        res = gqa.ood_evaluate((dset, loader, None), (z["maha.mean"], z["maha.precision"]), dump=dump)
    else:                      # gqa_conf.py:296 / gqa_energy.py:194: ood_evaluate(eval_tuple, dump=)
        res = gqa.ood_evaluate((dset, loader, None), dump=dump)
    want = meta["results"][method]
    assert set(res) == set(want)
    for k, v in want.items():
        assert float(res[k]) == v, (method, k, float(res[k]), v)
    rows = json.load(open(dump))
    assert [r["questionId"] for r in rows] == qids
    assert [r["prediction"] for r in rows] == [meta["label2ans"][int(i)] for i in z[f"{method}.label"]]
    np.testing.assert_allclose([r["confidence"] for r in rows], np.round(z[f"{method}.score"], 4), atol=1.01e-4)
    if method != "maha":       # Mahalanobis without the estimator's (mean, precision): the reference fails on the missing pickle too
        with pytest.raises(RuntimeError):
            gqa.ood_evaluate((dset, loader, None), method="maha")
