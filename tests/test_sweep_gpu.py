"""The public Python seam on the GPU: GQAModel.forward(feat, pos, sent), the GQAEval sweep (pinned host batches,
double-buffered H2D, device scorers) and the JSON it writes, against the CPU oracle run on the same inputs."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

from rgqa_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
LAYERS = dict(l_layers=2, x_layers=1, r_layers=1)


def _args(precision):
    return types.SimpleNamespace(llayers=2, xlayers=1, rlayers=1, from_scratch=False, precision=precision, vocab=None,
                                 max_seq_length=20)


@pytest.fixture(scope="module")
def small_model(cuda_device):
    sys.argv = ["x"]
    from rgqa_b200.tasks.gqa_model import GQAModel_maha
    m = GQAModel_maha(1842, args=_args("fp32"), max_seq_length=20)
    sd = syn.lxmert_state_dict(variant="rich", head_sigma=0.2, **LAYERS)
    m.load_state_dict(sd, strict=True)
    return m.to(cuda_device), sd


def test_model_forward_with_sentences_matches_oracle(small_model, cuda_device):
    from oracle import lxmert_oracle as O
    m, sd = small_model
    sents = ["Is the cat on the mat?", "what color is the large bus to the left of the tree", "", "Are there any zebras?"]
    inp = syn.make_inputs(len(sents), 20, seed=5)
    logits, pooled = m(inp["feat"].to(cuda_device), inp["pos"].to(cuda_device), sents)
    ids, mask, seg = m.lxrt_encoder.tokenizer.encode_batch(sents, 20)
    ref_logits, ref_pooled = O.gqa_model_forward(sd, inp["feat"], inp["pos"], torch.from_numpy(ids).long(),
                                                 torch.from_numpy(seg).long(), torch.from_numpy(mask).long())
    assert (pooled.cpu() - ref_pooled).abs().max() <= 1e-4
    assert (logits.cpu() - ref_logits).abs().max() <= 1e-3  # head02 logits reach +-30
    # the packed engine is rebuilt when a parameter changes
    with torch.no_grad():
        m.logit_fc._parameters["3|bias"].add_(1.0)
    logits2, _ = m(inp["feat"].to(cuda_device), inp["pos"].to(cuda_device), sents)
    torch.testing.assert_close(logits2, logits + 1.0, atol=1e-4, rtol=0)
    with torch.no_grad():
        m.logit_fc._parameters["3|bias"].sub_(1.0)


@pytest.mark.parametrize("method", ["msp", "energy", "maha"])
def test_sweep_matches_oracle_and_writes_reference_json(small_model, cuda_device, tmp_path, method):
    from oracle import lxmert_oracle as O
    from oracle import metrics_oracle as MO
    from oracle import scorers_oracle as S
    from rgqa_b200.tasks.gqa_eval import GQAEval, SyntheticRGQA
    m, sd = small_model
    ds = SyntheticRGQA(50, L=20, seed=11, pool=50)
    mu, P = syn.maha_stats()
    m.engine().set_maha(mu, P)
    dump = str(tmp_path / f"{method}_predict.json")
    res = GQAEval(m, ds, method=method, tau=0.5).ood_evaluate(ds.loader(16), dump=dump)  # ragged last batch (50 = 3*16+2)
    logit, x = O.gqa_model_forward(sd, ds.feat.float(), ds.pos, ds.ids.long(), torch.zeros_like(ds.ids).long(), ds.mask.long())
    rs, rl = {"msp": S.msp, "energy": S.energy_top2, "maha": lambda _: S.mahalanobis(x, mu, P)[:2]}[method](logit)
    rows = json.load(open(dump))
    assert [r["questionId"] for r in rows] == ds.qids
    assert [r["prediction"] for r in rows] == [ds.label2ans[int(i)] for i in rl]
    np.testing.assert_allclose([r["confidence"] for r in rows], np.round(rs.numpy().astype(np.float64), 4), rtol=1e-3, atol=2e-4)
    # metrics of OUR scores through the product evaluator == the sklearn oracle on the same triples (bit-exact)
    q2l = {q: ds.id2datum[q]["label"] for q in ds.qids}
    want = MO.evaluate(q2l, _triples(m, ds, method, cuda_device))
    for k, v in want.items():
        assert float(res[k]) == float(v), k


def _triples(m, ds, method, dev):
    from rgqa_b200.tasks.gqa_eval import SweepRunner
    r = SweepRunner(m, method, device=dev)
    out = []
    for qids, f, p, i, mk in ds.loader(16):
        s, l = r.submit(f, p, i, mk)
        out += [(q, ds.label2ans[int(a)], b) for q, a, b in zip(qids, l.cpu().numpy(), s.cpu().numpy())]
    return out


def test_uniter_module_matches_golden(cuda_device):
    sys.argv = ["x"]
    from rgqa_b200.uniter.uniter import GQAUNITER
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "uniter_rich_L20_B4.npz"))
    meta = json.loads(str(z["meta"]))
    m = GQAUNITER(1842, args=_args("fp32"))
    m.load_state_dict(syn.uniter_state_dict(variant="rich"), strict=True)
    m = m.to(cuda_device)
    inp = syn.make_inputs(4, 20, seed=meta["inputs_seed"], vocab=28996)
    logits, _ = m.forward_ids(inp["feat"].to(cuda_device), inp["pos7"].to(cuda_device), inp["ids"].to(cuda_device),
                              inp["seg"].to(cuda_device), inp["mask"].to(cuda_device))
    assert np.abs(logits.cpu().numpy() - z["logits"]).max() <= 1e-4


def test_bf16_feature_store_sweep_is_bit_identical_to_fp32_features(cuda_device, tmp_path):
    """features converted ONCE to a bf16 memory-mapped store (round-to-nearest-even on the host) give
    exactly the scores of fp32 features cast on the device — at half the H2D bytes.  TSV -> store -> SweepLoader ->
    SweepRunner, against the same questions pushed through the model from fp32 tensors."""
    import base64
    from rgqa_b200 import ingest
    from rgqa_b200.tasks.gqa_eval import SweepRunner
    from rgqa_b200.tasks.gqa_model import GQAModel_maha
    sys.argv = ["x"]
    m = GQAModel_maha(1842, args=_args("bf16"), max_seq_length=20)
    m.load_state_dict(syn.lxmert_state_dict(variant="rich", head_sigma=0.2, **LAYERS), strict=True)
    m = m.to(cuda_device)
    n_img = 5
    inp = syn.make_inputs(n_img, 20, seed=77)
    rng = np.random.default_rng(5)
    lines = []
    for i in range(n_img):
        w, h = 500, 400
        boxes = (inp["pos"][i].numpy() * np.array([w, h, w, h], dtype=np.float32)).astype(np.float32)
        vals = {"img_id": f"img{i}", "img_h": h, "img_w": w, "num_boxes": 36,
                "objects_id": rng.integers(0, 1600, 36).astype(np.int64), "objects_conf": rng.random(36).astype(np.float32),
                "attrs_id": rng.integers(0, 400, 36).astype(np.int64), "attrs_conf": rng.random(36).astype(np.float32),
                "boxes": boxes, "features": inp["feat"][i].numpy()}
        lines.append("\t".join(base64.b64encode(np.ascontiguousarray(vals[k]).tobytes()).decode() if isinstance(vals[k], np.ndarray)
                               else str(vals[k]) for k in ingest.FIELDNAMES))
    tsv = tmp_path / "obj.tsv"
    tsv.write_text("\n".join(lines) + "\n")
    store = ingest.convert_obj_tsv(str(tsv), str(tmp_path / "store"), dtype="bf16")
    sents = ["is the cat on the mat?", "what color is the bus", "are there zebras", "who holds the umbrella", "is it raining",
             "how many dogs are there", "what is left of the tree"]
    imgs = [f"img{i % n_img}" for i in range(len(sents))]
    data = [{"question_id": f"q{i}", "img_id": im, "sent": s} for i, (im, s) in enumerate(zip(imgs, sents))]
    tok = m.lxrt_encoder.tokenizer
    runner = SweepRunner(m, "energy")
    got_s, got_l, qids = [], [], []
    for q, f, b, ids, mask in ingest.SweepLoader(store, data, tok, max_seq_length=20, batch_size=4):
        assert f.dtype == torch.bfloat16 and f.is_pinned()
        s, l = runner.submit(f, b, ids, mask)
        got_s.append(s.cpu()); got_l.append(l.cpu()); qids += q
    assert qids == [d["question_id"] for d in data]
    # reference route: fp32 features and the reference's per-item box normalisation, straight to the model
    rows = [int(d["img_id"][3:]) for d in data]
    feat32 = inp["feat"][rows].to(cuda_device)
    boxes = torch.from_numpy(np.stack([ingest.normalize_boxes(inp["pos"][r].numpy() * np.array([500, 400, 500, 400], dtype=np.float32), 500, 400)
                                       for r in rows])).to(cuda_device)
    ids, mask, seg = tok.encode_batch(sents, 20)
    logits, pooled = m.forward_ids(feat32, boxes, torch.from_numpy(ids).to(cuda_device), torch.from_numpy(seg).to(cuda_device),
                                   torch.from_numpy(mask).to(cuda_device))
    ref_s, ref_l = m.engine().score("energy", logits=logits, pooled=pooled)
    assert torch.equal(torch.cat(got_s), ref_s.cpu()) and torch.equal(torch.cat(got_l), ref_l.cpu())


def test_gqa_train_mirror_runs_rp_steps_and_updates_the_model(cuda_device, tmp_path):
    """GQA.train (gqa_conf.py:140-202) through the host mirror: random-pairing batches, BCE x A, clip, BertAdam; the loss falls on
    a batch it has seen and the trained parameters land back in the model the evaluation path uses."""
    from rgqa_b200.tasks.gqa_conf import GQA
    sys.argv = ["x"]
    args = types.SimpleNamespace(backbone="lxmert", llayers=2, xlayers=1, rlayers=1, from_scratch=False, precision="fp32", vocab=None,
                                 max_seq_length=20, output=str(tmp_path / "out"),
                                 lr=2e-4, epochs=3, dropout=0.0, seed=9595, sample_pair=True, tau=0.5)
    gqa = GQA(args)
    gqa.model.load_state_dict(syn.lxmert_state_dict(variant="rich", **LAYERS), strict=True)
    gqa.model = gqa.model.to(cuda_device)
    B = 6
    inp = syn.make_inputs(B, 20, seed=41)
    sents = [f"what is item {i} doing near the tree" for i in range(B)]
    target = torch.zeros(B, 1843)
    target[torch.arange(B), torch.tensor([3, 50, 700, 11, 1500, 90])] = 1.0
    qids = [f"q{i}" for i in range(B)]
    loader = [(qids, inp["feat"], inp["pos"], sents, target)]
    before = gqa.model.state_dict()["logit_fc.3.bias"].clone()
    logits_before = gqa.model(inp["feat"].to(cuda_device), inp["pos"].to(cuda_device), sents)[0].clone()
    losses = gqa.train(loader, img_of=lambda q: "img" + q[1:], epochs=3)
    assert len(losses) == 3 and losses[2] < losses[0]
    after = gqa.model.state_dict()["logit_fc.3.bias"]
    assert not torch.equal(before.cpu(), after.cpu())
    logits_after = gqa.model(inp["feat"].to(cuda_device), inp["pos"].to(cuda_device), sents)[0]
    # the evaluation path sees the trained weights: its BCE x A loss on the training batch fell as well
    bce = torch.nn.BCEWithLogitsLoss()
    t = target[:, :-1].to(logits_after.device)
    assert float(bce(logits_after, t)) < float(bce(logits_before, t))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_drop_in_flow_reproduces_the_reference_stack_end_to_end(cuda_device, tmp_path, precision):
    """Sentences in -> ``<method>_predict.json`` + result dict out, against the UNMODIFIED reference stack run end to end in the
    build container (tests/golden/make_golden_dropin.py: the real ``GQAModel`` / ``GQAModel_maha`` built by their own constructors,
    the reference tokenizer, ``GQA.ood_evaluate`` of gqa_conf / gqa_energy / gqa_maha, the real evaluator's ``dump_result``).
    This side: ``GQAModel_maha.forward(feat, pos, sent)`` (same state dict, same vocab, MAX_GQA_LENGTH = 30), the device
    scorers, the product evaluator — the loop is the drivers' (gqa_conf.py:301-330)."""
    import gzip
    from rgqa_b200.tasks.gqa_data import GQAOODEvaluator
    from rgqa_b200.tasks.gqa_model import GQAModel_maha
    gold = os.path.join(os.path.dirname(__file__), "golden")
    d = json.load(gzip.open(os.path.join(gold, "dropin_reference.json.gz"), "rt"))
    sys.argv = ["x"]
    a = types.SimpleNamespace(llayers=d["layers"]["l_layers"], xlayers=d["layers"]["x_layers"], rlayers=d["layers"]["r_layers"],
                              from_scratch=False, precision=precision, vocab=os.path.join(gold, "tokenizer_vocab.txt"),
                              max_seq_length=d["max_seq_length"])
    m = GQAModel_maha(d["num_answers"], args=a, max_seq_length=d["max_seq_length"])
    m.load_state_dict(syn.lxmert_state_dict(variant="rich", head_sigma=d["head_sigma"], **d["layers"]), strict=True)
    m = m.to(cuda_device)
    mu, P = syn.maha_stats()
    m.engine().set_maha(mu, P)
    inp = syn.make_inputs(len(d["qids"]), 20, seed=d["inputs_seed"])
    label2ans = [f"answer{j}" for j in range(d["num_answers"])] + ["UQ"]
    qid2label = {q: dict(map(tuple, l)) for q, l in zip(d["qids"], d["labels"])}
    bs = d["batch"]
    # bf16 operands move the logits by ~1e-2: MSP sits on the flat part of the sigmoid, energy scales with the logits,
    # Mahalanobis (|score| ~ 400 here) is held to a RELATIVE bound
    tol = {"fp32": {"msp": 1.5e-4, "energy": 2e-4, "maha": 2e-3}, "bf16": {"msp": 1.5e-3, "energy": 6e-2, "maha": 5e-3}}[precision]
    ref32 = None
    if precision == "bf16":   # the fp32 CUDA path on the same inputs gives the margins that decide which label swaps are legitimate
        a32 = types.SimpleNamespace(**{**vars(a), "precision": "fp32"})
        m32 = GQAModel_maha(d["num_answers"], args=a32, max_seq_length=d["max_seq_length"])
        m32.load_state_dict(syn.lxmert_state_dict(variant="rich", head_sigma=d["head_sigma"], **d["layers"]), strict=True)
        m32 = m32.to(cuda_device)
        m32.engine().set_maha(mu, P)
        lg, xs, full = [], [], []
        for lo in range(0, len(d["qids"]), bs):
            sl = slice(lo, lo + bs)
            logit, x = m32(inp["feat"][sl].to(cuda_device), inp["pos"][sl].to(cuda_device), d["sents"][sl])
            lg.append(logit)
            full.append(m32.engine().score("maha", pooled=x, return_full=True)[2])
        lg, full = torch.cat(lg), torch.cat(full)
        t2l, t2m = lg.topk(2, dim=1).values, full.topk(2, dim=1).values
        # a label may differ from the reference's only where the reference's own top-2 margin is inside the bf16 noise:
        # 2 x the 2e-2 logit tolerance for the logit scorers, 2 x 1e-3 relative for the Mahalanobis score
        ref32 = {"msp": (t2l[:, 0] - t2l[:, 1]).cpu() > 4e-2, "energy": (t2l[:, 0] - t2l[:, 1]).cpu() > 4e-2,
                 "maha": ((t2m[:, 0] - t2m[:, 1]) > 2e-3 * t2m[:, 0].abs()).cpu()}
    for method in ("msp", "energy", "maha"):
        quesid2ans = {}
        for lo in range(0, len(d["qids"]), bs):
            sl = slice(lo, lo + bs)
            logit, x = m(inp["feat"][sl].to(cuda_device), inp["pos"][sl].to(cuda_device), d["sents"][sl])
            score, label = m.engine().score(method, logits=logit, pooled=x)
            for qid, l, s in zip(d["qids"][sl], label.cpu().numpy(), score.cpu().numpy()):
                quesid2ans[qid] = (label2ans[l], s)
        ev = GQAOODEvaluator(qid2label, 0.5)
        res = ev.evaluate(quesid2ans)
        path = str(tmp_path / f"{method}_predict.json")
        ev.dump_result(quesid2ans, path)
        rows, want = json.load(open(path)), json.loads(d["predict_json"][method])
        assert [r["questionId"] for r in rows] == [r["questionId"] for r in want]
        same = [r["prediction"] == w["prediction"] for r, w in zip(rows, want)]
        if precision == "fp32":
            assert all(same), (method, same)
        else:
            # labels of near-tied classes may swap under bf16's logit noise — but ONLY those: every question whose
            # fp32 margin is clear of the tolerance must carry the reference's answer
            clear = ref32[method]
            assert all(ok for ok, c in zip(same, clear.tolist()) if c), (method, [i for i, (ok, c) in enumerate(zip(same, clear.tolist())) if c and not ok])
            assert int(clear.sum()) >= len(same) // 4, (method, int(clear.sum()))   # the check is not vacuous
        if precision == "bf16" and method == "maha":   # relative: |score| ~ 400
            conf_err = max(abs(r["confidence"] - w["confidence"]) / abs(w["confidence"]) for r, w, ok in zip(rows, want, same) if ok)
        else:
            conf_err = max(abs(r["confidence"] - w["confidence"]) for r, w, ok in zip(rows, want, same) if ok)
        assert conf_err <= tol[method], (method, conf_err)
        if precision == "fp32":
            assert set(res) == set(d["results"][method])
            for k, v in d["results"][method].items():
                assert abs(float(res[k]) - v) <= 1e-6 + 2e-2 * abs(v), (method, k, float(res[k]), v)
            if all(r["confidence"] == w["confidence"] for r, w in zip(rows, want)):
                assert open(path).read() == d["predict_json"][method]          # byte-identical wire file


def test_uniter_drop_in_flow_reproduces_the_reference_stack_end_to_end(cuda_device, tmp_path):
    """BASELINE config 3: ``GQAUNITER.forward(feat, pos7, sent)`` + device scorers + product evaluator against the UNMODIFIED
    reference ``GQAUNITER`` (12 layers, own constructor and tokenizer) driven by ``gqa_conf`` / ``gqa_energy`` ``GQA.ood_evaluate``
    (tests/golden/make_golden_dropin_uniter.py), fp32 path."""
    import gzip
    from rgqa_b200.tasks.gqa_data import GQAOODEvaluator
    from rgqa_b200.uniter.uniter import GQAUNITER
    gold = os.path.join(os.path.dirname(__file__), "golden")
    d = json.load(gzip.open(os.path.join(gold, "dropin_uniter_reference.json.gz"), "rt"))
    sys.argv = ["x"]
    a = _args("fp32")
    a.vocab = os.path.join(gold, "tokenizer_vocab.txt")
    m = GQAUNITER(d["num_answers"], args=a)
    m.load_state_dict(syn.uniter_state_dict(variant="rich", head_sigma=d["head_sigma"]), strict=True)
    m = m.to(cuda_device)
    inp = syn.make_inputs(len(d["qids"]), 20, seed=d["inputs_seed"], vocab=28996)
    label2ans = [f"answer{j}" for j in range(d["num_answers"])] + ["UQ"]
    qid2label = {q: dict(map(tuple, l)) for q, l in zip(d["qids"], d["labels"])}
    bs = d["batch"]
    for method in ("msp", "energy"):
        quesid2ans = {}
        for lo in range(0, len(d["qids"]), bs):
            sl = slice(lo, lo + bs)
            logit = m(inp["feat"][sl].to(cuda_device), inp["pos7"][sl].to(cuda_device), d["sents"][sl])
            score, label = m.engine().score(method, logits=logit)
            for qid, l, s in zip(d["qids"][sl], label.cpu().numpy(), score.cpu().numpy()):
                quesid2ans[qid] = (label2ans[l], s)
        ev = GQAOODEvaluator(qid2label, 0.5)
        res = ev.evaluate(quesid2ans)
        path = str(tmp_path / f"{method}_predict.json")
        ev.dump_result(quesid2ans, path)
        rows, want = json.load(open(path)), json.loads(d["predict_json"][method])
        assert [(r["questionId"], r["prediction"]) for r in rows] == [(w["questionId"], w["prediction"]) for w in want]
        assert max(abs(r["confidence"] - w["confidence"]) for r, w in zip(rows, want)) <= 2e-4
        for k, v in d["results"][method].items():
            assert abs(float(res[k]) - v) <= 1e-6 + 2e-2 * abs(v), (method, k, float(res[k]), v)
