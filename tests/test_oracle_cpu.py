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
