"""Host logic of the RP-finetune groundwork: learning-rate schedules against the values the
reference optimiser reported (tests/golden/bertadam_steps.npz, made by make_golden_bertadam.py) and the random-pairing
batch builder against src/tasks/gqa_conf.py:155-170 executed as written."""
import os
import random

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_schedules_match_reference_values():
    from rgqa_b200.lxrt.optimization import SCHEDULES, warmup_constant, warmup_linear
    z = np.load(os.path.join(GOLD, "bertadam_steps.npz"))
    for k in range(4):
        assert 1e-3 * warmup_linear(k / 10, 0.1) == pytest.approx(float(z[f"lr{k}"]), rel=1e-12, abs=0)
    assert warmup_linear(1.5, 0.1) == 0 and warmup_constant(0.001) == 0.5 and warmup_constant(0.5) == 1.0
    assert set(SCHEDULES) == {"warmup_cosine", "warmup_constant", "warmup_linear"}
    assert SCHEDULES["warmup_cosine"](0.5, 0.1) == pytest.approx(0.5)


def test_rp_batch_matches_reference_lines():
    from rgqa_b200.tasks.rp import build_rp_batch
    B = 9
    ques_id = [f"q{i}" for i in range(B)]
    img = {q: f"img{i // 2}" for i, q in enumerate(ques_id)}            # pairs of questions share an image
    id2datum = {q: {"img_id": img[q]} for q in ques_id}
    sent = [f"sentence {i}" for i in range(B)]
    feats, boxes = torch.randn(B, 36, 8), torch.rand(B, 36, 4)
    target = (torch.rand(B, 11) < 0.2).float()
    # the reference lines, executed as written (gqa_conf.py:155-170)
    random.seed(123)
    neg_sent, neg_qid = [], []
    for j, qid, in enumerate(ques_id):
        qid_r = random.choice(ques_id)
        while id2datum[qid_r]['img_id'] == id2datum[qid]['img_id']:
            qid_r = random.choice(ques_id)
        ind_r = ques_id.index(qid_r)
        neg_sent.append(sent[ind_r])
        neg_qid.append(id2datum[qid]['img_id'] + '+' + qid_r)
    want = (ques_id + neg_qid, feats.repeat(2, 1, 1), boxes.repeat(2, 1, 1), sent + neg_sent,
            torch.cat([target, torch.zeros_like(target)], 0))
    random.seed(123)
    got = build_rp_batch(ques_id, feats, boxes, sent, target, lambda q: id2datum[q]["img_id"])
    assert got[0] == want[0] and got[3] == want[3]
    assert torch.equal(got[1], want[1]) and torch.equal(got[2], want[2]) and torch.equal(got[4], want[4])
    for q, nq in zip(ques_id, got[0][B:]):                                # a negative never keeps the row's own image
        assert img[nq.split("+")[1]] != img[q] and nq.startswith(img[q] + "+")
    with pytest.raises(ValueError):
        build_rp_batch(["a", "b"], feats[:2], boxes[:2], sent[:2], target[:2], lambda q: "same")


def _rp_golden():
    import json
    gold = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(gold, "rp_train_reference.npz"))
    with open(os.path.join(gold, "rp_train_reference.json")) as f:
        meta = json.load(f)
    return z, meta


def test_rp_batches_equal_what_the_reference_train_loop_fed_its_model():
    """``build_rp_batch`` against the UNMODIFIED ``GQA.train`` (src/tasks/gqa_conf.py:148-170, run by
    tests/golden/make_golden_rp_train.py): same ``random`` seed -> the same sentences, in the same order, for every step of two
    epochs — i.e. the same draws, the same redraws while the drawn question shares the row's image, the same batch layout."""
    from rgqa_b200.tasks.rp import build_rp_batch
    z, meta = _rp_golden()
    qids, sents, img = meta["qids"], meta["sents"], meta["img_of"]
    feats, boxes, target = (torch.from_numpy(z[k]) for k in ("feat", "box", "target"))
    bs, step = meta["batch"], 0
    random.seed(meta["seed"])
    for _ in range(meta["epochs"]):
        for lo in range(0, len(qids), bs):
            sl = slice(lo, lo + bs)
            q2, f2, b2, s2, t2 = build_rp_batch(qids[sl], feats[sl], boxes[sl], sents[sl], target[sl][:, :-1], img.__getitem__)
            want = meta["steps"][step]
            assert s2 == want["sent"], step
            assert f2.shape[0] == b2.shape[0] == t2.shape[0] == len(q2) == want["rows"]
            assert int((t2.abs().sum(1) > 0).sum()) == want["target_rows_nonzero"]
            assert torch.equal(f2[:bs], f2[bs:]) and not t2[bs:].any()
            step += 1
    assert step == len(meta["steps"])
