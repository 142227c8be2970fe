"""Pin the CPU oracle against golden vectors produced by the UNMODIFIED reference modules
(tests/golden/make_golden.py, run in the build container).  fp32 on both sides: only GEMM summation
order differs, so the tolerance is 2e-5 abs on O(1) activations / logits."""
import json
import os

import numpy as np
import pytest
import torch

from rgqa_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(tag):
    z = np.load(os.path.join(GOLD, tag + ".npz"))
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("tag", ["lxmert_init_L20_B4", "lxmert_rich_L30_B4", "lxmert_rich_L20_B6_head02"])
def test_lxmert_oracle_matches_reference_golden(tag):
    from oracle import lxmert_oracle as O
    z, meta = _load(tag)
    sd = syn.lxmert_state_dict(seed=meta["weights_seed"], variant=meta["variant"], head_sigma=meta["head_sigma"])
    inp = syn.make_inputs(meta["B"], meta["L"], seed=meta["inputs_seed"])
    with torch.no_grad():
        pooled, trace = O.lxmert_encode(sd, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], return_trace=True)
        logits = O.gqa_head(sd, pooled)
    np.testing.assert_allclose(trace["emb"][0].numpy(), z["emb0"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(trace["visn_fc"][0].numpy(), z["vfc0"], atol=2e-5, rtol=0)
    last = max(int(k[6:]) for k in trace if k.startswith("lang_x"))
    np.testing.assert_allclose(trace[f"lang_x{last}"][0].numpy(), z["lang0"], atol=5e-5, rtol=0)
    np.testing.assert_allclose(trace[f"visn_x{last}"][0].numpy(), z["visn0"], atol=5e-5, rtol=0)
    np.testing.assert_allclose(pooled.numpy(), z["pooled"], atol=2e-5, rtol=0)
    scale = max(1.0, float(np.abs(z["logits"]).max()))
    np.testing.assert_allclose(logits.numpy(), z["logits"], atol=5e-5 * scale, rtol=0)
    assert (logits.argmax(1).numpy() == z["logits"].argmax(1)).all()


@pytest.mark.parametrize("tag", ["uniter_init_L20_B4", "uniter_rich_L20_B4"])
def test_uniter_oracle_matches_reference_golden(tag):
    from oracle import lxmert_oracle as O
    z, meta = _load(tag)
    sd = syn.uniter_state_dict(seed=meta["weights_seed"], variant=meta["variant"])
    inp = syn.make_inputs(meta["B"], meta["L"], seed=meta["inputs_seed"], vocab=28996)
    logits, pooled = O.gqa_uniter_forward(sd, inp["feat"], inp["pos7"], inp["ids"], inp["seg"], inp["mask"])
    np.testing.assert_allclose(pooled.numpy(), z["pooled"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(logits.numpy(), z["logits"], atol=5e-5, rtol=0)


def test_sample_estimator_matches_sklearn_empirical_covariance():
    """The reference calls sklearn.covariance.EmpiricalCovariance().fit(X).precision_ (gqa_maha.py:129-183)."""
    import sklearn.covariance
    from oracle import scorers_oracle as S
    g = torch.Generator().manual_seed(1)
    N, C, D = 400, 7, 24
    feats = torch.randn(N, D, generator=g) + torch.arange(D) * 0.1
    targets = (torch.rand(N, C, generator=g) < 0.25).float()
    targets[torch.arange(N), torch.randint(0, C - 1, (N,), generator=g)] = 1.0  # last class may stay empty
    targets[:, C - 1] = 0
    mean, prec = S.sample_estimator(feats, targets)
    # literal restatement with sklearn
    X = []
    for j in range(C):
        idx = targets[:, j] != 0
        if idx.any():
            assert torch.allclose(mean[j], feats[idx].mean(0), atol=1e-6)
            X.append(feats[idx] - feats[idx].mean(0))
        else:
            assert torch.equal(mean[j], torch.zeros(D))
    ref = sklearn.covariance.EmpiricalCovariance(assume_centered=False).fit(torch.cat(X, 0).numpy()).precision_
    np.testing.assert_allclose(prec, ref, rtol=1e-6, atol=1e-8)


def test_mahalanobis_oracle_equals_expanded_form():
    """The GEMM form the CUDA kernel uses is algebraically the reference loop (checked here in float64)."""
    from oracle import scorers_oracle as S
    mu, P = syn.maha_stats(num_answers=50, hidden=64, n=512)
    g = torch.Generator().manual_seed(3)
    x = torch.tanh(torch.randn(5, 64, generator=g))
    score, label, full = S.mahalanobis(x.double(), mu.double(), P.double())
    t = x.double() @ P.double()
    q = (t * x.double()).sum(1, keepdim=True)
    G = t @ mu.double().t()
    m = ((mu.double() @ P.double()) * mu.double()).sum(1)
    torch.testing.assert_close(full, -0.5 * (q - 2 * G + m), rtol=1e-10, atol=1e-10)
    assert torch.equal(label, full.argmax(1))


def test_ensemble_oracle_matches_the_reference_loop_literally():
    """src/tasks/gqa_ensemble.py:165-177 executed as written (dict of per-question float32 vectors, in-place += / *=,
    / len for the mean, np.argmax) against the vectorised oracle."""
    from oracle import scorers_oracle as S
    rng = np.random.default_rng(7)
    n, A, K = 37, 1842, 3
    per_model = [1.0 / (1.0 + np.exp(-rng.normal(0, 2.0, size=(n, A)))).astype(np.float32) for _ in range(K)]
    per_model[1][5] = per_model[0][5]            # ties / identical rows are legal inputs
    for method in ("mean", "multiply"):
        results_list = [{q: per_model[k][q].copy() for q in range(n)} for k in range(K)]
        final_results = results_list[0]
        for results in results_list[1:]:
            for qid in results:
                if method == "mean":
                    final_results[qid] += results[qid]
                else:
                    final_results[qid] *= results[qid]
        want = {}
        for qid in final_results:
            scores = final_results[qid] / len(results_list) if method == "mean" else final_results[qid]
            label = np.argmax(scores)
            want[qid] = (label, scores[label])
        score, label = S.ensemble(per_model, method)
        assert [int(l) for l in label] == [int(want[q][0]) for q in range(n)]
        assert np.array_equal(score, np.array([want[q][1] for q in range(n)], dtype=np.float32))
    with pytest.raises(ValueError):
        S.ensemble(per_model, "median")


def test_estimator_closed_form_finish_matches_the_list_building_oracle():
    """rgqa_b200.tasks.gqa_maha.finish_estimator turns streamed sums (n_c, s_c, G = sum w_i x x^T) into the reference's
    (class means, precision) pair; here the sums are formed with numpy and the result compared with the oracle's literal
    restatement of src/tasks/gqa_maha.py:150-183 (multi-label rows and empty classes included)."""
    from oracle import scorers_oracle as S
    from rgqa_b200.tasks.gqa_maha import finish_estimator
    g = torch.Generator().manual_seed(1)
    N, C, D = 600, 30, 64
    x = torch.tanh(torch.randn(N, D, generator=g))
    t = torch.zeros(N, C)
    t[torch.arange(N), torch.randint(0, C - 2, (N,), generator=g)] = 1
    t[5, :10] = 1
    ref_mean, ref_prec = S.sample_estimator(x, t)
    xd, td = x.double().numpy(), t.double().numpy()
    w = td.sum(1)
    mean, prec = finish_estimator((xd * w[:, None]).T @ xd, td.T @ xd, td.sum(0))
    assert (mean - ref_mean).abs().max().item() <= 1e-6
    assert torch.equal(mean[C - 2:], torch.zeros(2, D))
    assert np.abs(prec - ref_prec).max() <= 1e-6 * np.abs(ref_prec).max()


# ------------------------------------------------------------------------------------------------------------------
# Scorers pinned to the UNMODIFIED reference drivers (tests/golden/make_golden_scorers.py ran gqa_conf / gqa_energy /
# gqa_odin / gqa_maha / gqa_ensemble `GQA.ood_evaluate` + `sample_estimator` on the tiny model of tests/golden/tiny_vqa.py)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def scorer_golden():
    import sys
    sys.path.insert(0, GOLD)
    from tiny_vqa import TinyVQA
    z = np.load(os.path.join(GOLD, "scorers_reference.npz"))
    with open(os.path.join(GOLD, "scorers_reference.json")) as f:
        meta = json.load(f)
    models = []
    for tag in ("model", "model2"):
        sd = {k[len(tag) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + ".")}
        C, D = sd["head.weight"].shape
        m = TinyVQA(sd["fc_feat.weight"].shape[1], D, C, maha=True)
        m.load_state_dict(sd, strict=True)
        models.append(m.eval())
    return z, meta, models


def _batches(n, bs=16):   # the reference loader's batches (ODIN's loss is a batch mean: keep them identical)
    return [slice(i, min(i + bs, n)) for i in range(0, n, bs)]


def test_scorer_oracles_pinned_to_reference_drivers(scorer_golden):
    from oracle import scorers_oracle as S
    from oracle import odin_oracle as OD
    z, meta, (model, model2) = scorer_golden
    feat, box = torch.from_numpy(z["feat_test"]), torch.from_numpy(z["box_test"])
    sents = meta["sents_test"]
    n = feat.shape[0]
    with torch.no_grad():
        logit, x = model(feat, box, sents)
        logit2, _ = model2(feat, box, sents)

    # MSP (gqa_conf.py:326) and top-2 energy (gqa_energy.py:204-206): same floats, same labels
    for name, fn in (("msp", S.msp), ("energy", S.energy_top2)):
        s, l = fn(logit)
        np.testing.assert_array_equal(l.numpy(), z[f"{name}.label"])
        np.testing.assert_allclose(s.double().numpy(), z[f"{name}.score"], rtol=1e-6, atol=0)

    # the two training-time energies, locals of the reference `GQA.train` loop (gqa_energy.py:131 with --mceLoss, :135)
    for name, fn in (("train_energy_sum", S.energy_sum), ("train_lse", S.lse)):
        e, _ = fn(torch.from_numpy(z[f"{name}.logit"]))
        np.testing.assert_allclose(e.numpy(), z[f"{name}.E_f"], rtol=1e-6)

    # ensemble members (gqa_ensemble.py:95-110): the per-question sigmoid vectors the combine step consumes
    np.testing.assert_allclose(torch.sigmoid(logit).numpy(), z["ens0.sigmoid"], rtol=1e-6)
    np.testing.assert_allclose(torch.sigmoid(logit2).numpy(), z["ens1.sigmoid"], rtol=1e-6)

    # ... and the combine step itself (gqa_ensemble.py:155-177, executed from the reference script's __main__ block)
    for method in ("mean", "multiply"):
        es, el = S.ensemble([z["ens0.sigmoid"], z["ens1.sigmoid"]], method)
        np.testing.assert_array_equal(el, z[f"ens_{method}.label"])
        np.testing.assert_allclose(es.astype(np.float64), z[f"ens_{method}.score"], rtol=1e-6)
    assert np.abs(z["ens_mean.score"] - z["ens_multiply.score"]).max() > 1e-3

    # sample_estimator (gqa_maha.py:120-189): class means + sklearn EmpiricalCovariance.precision_
    ftr, btr = torch.from_numpy(z["feat_train"]), torch.from_numpy(z["box_train"])
    with torch.no_grad():
        _, xtr = model(ftr, btr, meta["sents_train"])
    mean, precision = S.sample_estimator(xtr, torch.from_numpy(z["target_train"])[:, :-1])
    np.testing.assert_allclose(mean.numpy(), z["maha.mean"], rtol=1e-6, atol=1e-7)
    scale = np.abs(z["maha.precision"]).max()
    np.testing.assert_allclose(precision, z["maha.precision"], rtol=1e-7, atol=1e-9 * scale)

    # Mahalanobis score, noise = 0 (gqa_maha.py:228-236, :270) with the reference's own (mean, precision)
    mu, P = torch.from_numpy(z["maha.mean"]).float(), torch.from_numpy(z["maha.precision"]).float()
    s, l, _ = S.mahalanobis(x, mu, P)
    np.testing.assert_array_equal(l.numpy(), z["maha.label"])
    np.testing.assert_allclose(s.double().numpy(), z["maha.score"], rtol=2e-5)

    # the two input-gradient scorers, batch by batch like the reference loader
    def fwd(sl):
        return lambda f, b: model(f, b, sents[sl])
    got = {k: ([], []) for k in ("odin", "odin_plain", "maha_noise")}
    for sl in _batches(n):
        for key, T, noise in (("odin", meta["temperature"], meta["noise"]), ("odin_plain", 1.0, 0.0)):
            s, l, _, _ = OD.odin(None, feat[sl], box[sl], None, None, None, T, noise, model=fwd(sl))
            got[key][0].append(s), got[key][1].append(l)
        s, l, _, _ = OD.mahalanobis_with_noise(None, feat[sl], box[sl], None, None, None, mu, P, meta["noise"], model=fwd(sl))
        got["maha_noise"][0].append(s), got["maha_noise"][1].append(l)
    for key, (ss, ll) in got.items():
        np.testing.assert_array_equal(torch.cat(ll).numpy(), z[f"{key}.label"], err_msg=key)
        np.testing.assert_allclose(torch.cat(ss).double().numpy(), z[f"{key}.score"], rtol=2e-5, err_msg=key)
    # the perturbation does something on this model (otherwise the noise > 0 goldens would pin nothing)
    assert np.abs(z["maha_noise.score"] - z["maha.score"]).max() > 1e-3
    assert np.abs(z["odin.score"] - z["odin_plain.score"]).max() > 1e-3


@pytest.mark.parametrize("name", ["msp", "energy", "odin", "maha", "maha_noise", "ens_mean", "ens_multiply"])
def test_evaluator_reproduces_reference_result_dicts_from_driver_scores(scorer_golden, name):
    """Product host code (no GPU needed): the reference drivers' (answer, score) pairs through this package's evaluator give the
    result dict the reference's ``GQAOODEvaluator`` printed for them, float for float."""
    from rgqa_b200.tasks.gqa_data import GQAOODEvaluator
    z, meta, _ = scorer_golden
    qid2label = {q: {a: v for a, v in pairs} for q, pairs in zip(meta["qids_test"], meta["labels_test"])}
    # the drivers hand the evaluator numpy float32 scalars (`score.cpu().numpy()` elements)
    q2a = {q: (meta["label2ans"][int(l)], np.float32(s))
           for q, l, s in zip(meta["qids_test"], z[f"{name}.label"], z[f"{name}.score"])}
    got = GQAOODEvaluator(qid2label, 0.5).evaluate(q2a)
    want = meta["results"][name]
    assert set(got) == set(want)
    for k in want:
        assert float(got[k]) == want[k], (name, k, float(got[k]), want[k])


def test_oracle_and_host_code_reproduce_the_reference_stack_wire_files(tmp_path):
    """CPU end to end: product tokenizer -> oracle LXMERT + scorers -> product evaluator / ``dump_result`` against the UNMODIFIED
    reference stack (real ``GQAModel``, reference tokenizer, ``GQA.ood_evaluate`` of gqa_conf / gqa_energy / gqa_maha, real
    ``GQAOODEvaluator``; tests/golden/make_golden_dropin.py).  MSP / energy: the ``*_predict.json`` files are byte-identical; Mahalanobis:
    same answers, confidences within one unit of the 4th decimal; all three result dicts float for float."""
    import gzip
    from oracle import lxmert_oracle as O
    from oracle import scorers_oracle as S
    from rgqa_b200.lxrt.tokenization import BertTokenizer
    from rgqa_b200.tasks.gqa_data import GQAOODEvaluator
    d = json.load(gzip.open(os.path.join(GOLD, "dropin_reference.json.gz"), "rt"))
    tok = BertTokenizer.from_pretrained(os.path.join(GOLD, "tokenizer_vocab.txt"), do_lower_case=True)
    sd = syn.lxmert_state_dict(variant="rich", head_sigma=d["head_sigma"], **d["layers"])
    inp = syn.make_inputs(len(d["qids"]), 20, seed=d["inputs_seed"])
    ids, mask, seg = tok.encode_batch(d["sents"], d["max_seq_length"])
    logit, x = O.gqa_model_forward(sd, inp["feat"], inp["pos"], torch.from_numpy(ids).long(), torch.from_numpy(seg).long(),
                                   torch.from_numpy(mask).long())
    mu, P = syn.maha_stats()
    label2ans = [f"answer{j}" for j in range(d["num_answers"])] + ["UQ"]
    qid2label = {q: dict(map(tuple, l)) for q, l in zip(d["qids"], d["labels"])}
    for method, fn in (("msp", S.msp), ("energy", S.energy_top2), ("maha", lambda _: S.mahalanobis(x, mu, P)[:2])):
        s, l = fn(logit)
        q2a = {q: (label2ans[int(a)], b) for q, a, b in zip(d["qids"], l.numpy(), s.numpy())}
        ev = GQAOODEvaluator(qid2label, 0.5)
        res = ev.evaluate(q2a)
        path = str(tmp_path / f"{method}_predict.json")
        ev.dump_result(q2a, path)
        rows, want = json.load(open(path)), json.loads(d["predict_json"][method])
        assert [(r["questionId"], r["prediction"]) for r in rows] == [(w["questionId"], w["prediction"]) for w in want]
        assert max(abs(r["confidence"] - w["confidence"]) for r, w in zip(rows, want)) <= (1.01e-4 if method == "maha" else 0.0)
        if method != "maha":
            assert open(path).read() == d["predict_json"][method]
        assert set(res) == set(d["results"][method])
        for k, v in d["results"][method].items():
            assert float(res[k]) == v, (method, k)


def test_uniter_oracle_and_host_code_reproduce_the_reference_stack_wire_files(tmp_path):
    """BASELINE config 3 on the CPU: product tokenizer -> oracle UNITER + scorers -> product evaluator against the UNMODIFIED
    reference ``GQAUNITER`` driven by ``gqa_conf`` / ``gqa_energy`` ``GQA.ood_evaluate`` (tests/golden/make_golden_dropin_uniter.py)."""
    import gzip
    from oracle import lxmert_oracle as O
    from oracle import scorers_oracle as S
    from rgqa_b200.lxrt.tokenization import BertTokenizer
    from rgqa_b200.tasks.gqa_data import GQAOODEvaluator
    d = json.load(gzip.open(os.path.join(GOLD, "dropin_uniter_reference.json.gz"), "rt"))
    tok = BertTokenizer.from_pretrained(os.path.join(GOLD, "tokenizer_vocab.txt"), do_lower_case=True)
    sd = syn.uniter_state_dict(variant="rich", head_sigma=d["head_sigma"])
    inp = syn.make_inputs(len(d["qids"]), 20, seed=d["inputs_seed"], vocab=28996)
    ids, mask, seg = tok.encode_batch(d["sents"], d["max_seq_length"])
    logit, _ = O.gqa_uniter_forward(sd, inp["feat"], inp["pos7"], torch.from_numpy(ids).long(), torch.from_numpy(seg).long(),
                                    torch.from_numpy(mask).long())
    label2ans = [f"answer{j}" for j in range(d["num_answers"])] + ["UQ"]
    qid2label = {q: dict(map(tuple, l)) for q, l in zip(d["qids"], d["labels"])}
    for method, fn in (("msp", S.msp), ("energy", S.energy_top2)):
        s, l = fn(logit)
        q2a = {q: (label2ans[int(a)], b) for q, a, b in zip(d["qids"], l.numpy(), s.numpy())}
        ev = GQAOODEvaluator(qid2label, 0.5)
        res = ev.evaluate(q2a)
        path = str(tmp_path / f"{method}_predict.json")
        ev.dump_result(q2a, path)
        rows, want = json.load(open(path)), json.loads(d["predict_json"][method])
        assert [(r["questionId"], r["prediction"]) for r in rows] == [(w["questionId"], w["prediction"]) for w in want]
        assert max(abs(r["confidence"] - w["confidence"]) for r, w in zip(rows, want)) <= 1.01e-4
        for k, v in d["results"][method].items():
            assert abs(float(res[k]) - v) <= 1e-12, (method, k)
