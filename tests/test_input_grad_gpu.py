"""SURVEY §8(f) rank 1 on the GPU: the taped fp32 forward, d loss / d (feats, boxes) against torch autograd through the
CPU oracle, and the two scorers built on it (ODIN, Mahalanobis with noise) against the oracle's restatement of the
reference lines."""
import numpy as np
import pytest
import torch

from rgqa_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
LAYERS = dict(l_layers=2, x_layers=2, r_layers=1)


@pytest.fixture(scope="module")
def setup(cuda_device):
    from rgqa_b200.lxrt.input_grad import LXMERTInputGrad
    sd = syn.lxmert_state_dict(variant="rich", **LAYERS)      # default head: max sigmoid ~0.9, so the BCE gradient is alive
    ig = LXMERTInputGrad(sd, device=cuda_device, **LAYERS)
    inp = syn.make_inputs(3, 20, seed=21)
    return sd, ig, inp


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def test_taped_forward_equals_engine_and_oracle(setup, cuda_device):
    from oracle import lxmert_oracle as O
    from rgqa_b200.engine import Engine, EngineConfig
    sd, ig, inp = setup
    logits, pooled = ig.forward(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"])
    ref_logits, ref_pooled = O.gqa_model_forward(sd, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"])
    assert (pooled.cpu() - ref_pooled).abs().max() <= 1e-4 and (logits.cpu() - ref_logits).abs().max() <= 1e-3
    eng = Engine(EngineConfig(precision="fp32", **LAYERS), device=str(cuda_device))
    eng.load_weights(sd.items())
    d = eng.device
    p2, l2 = eng.forward(*(inp[k].to(d) for k in ("feat", "pos", "ids", "seg", "mask")))
    assert (l2 - logits).abs().max().item() <= 1e-4        # same kernels (GELU as a separate pass here)
    eng.close()


@pytest.mark.parametrize("what", ["logits", "pooled", "both"])
def test_input_gradients_match_autograd_through_the_oracle(setup, what):
    from oracle import lxmert_oracle as O
    sd, ig, inp = setup
    g = torch.Generator().manual_seed(5)
    feats, boxes = inp["feat"].clone().requires_grad_(True), inp["pos"].clone().requires_grad_(True)
    x = O.lxmert_encode(sd, feats, boxes, inp["ids"], inp["seg"], inp["mask"])
    logit = O.gqa_head(sd, x)
    wl, wp = torch.randn(logit.shape, generator=g), torch.randn(x.shape, generator=g)
    loss = (logit * wl).sum() * (what != "pooled") + (x * wp).sum() * (what != "logits")
    loss.backward()
    ig.forward(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"])
    dfeat, dpos = ig.backward(dlogits=None if what == "pooled" else wl, dpooled=None if what == "logits" else wp)
    assert _rel(dfeat.cpu(), feats.grad) <= 2e-4, _rel(dfeat.cpu(), feats.grad)
    assert _rel(dpos.cpu(), boxes.grad) <= 2e-4, _rel(dpos.cpu(), boxes.grad)
    big = feats.grad.abs() > 1e-3 * feats.grad.abs().max()
    assert (torch.sign(dfeat.cpu())[big] == torch.sign(feats.grad)[big]).float().mean().item() >= 0.9999


def test_odin_scores_match_oracle(setup):
    from oracle import odin_oracle as OO
    from rgqa_b200.tasks.gqa_odin import odin_scores
    sd, ig, inp = setup
    for T, noise in ((1.0, 1e-4), (2.0, 1e-3)):
        ref_s, ref_l, ref_gf, ref_gb = OO.odin(sd, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], T, noise)
        s, l, gf, gb = odin_scores(ig, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], T, noise)
        assert _rel(gf.cpu(), ref_gf) <= 2e-4 and _rel(gb.cpu(), ref_gb) <= 2e-4
        assert torch.equal(l.cpu().long(), ref_l)
        torch.testing.assert_close(s.cpu(), ref_s, rtol=1e-3, atol=1e-6)


def test_mahalanobis_with_noise_matches_oracle(setup):
    from oracle import odin_oracle as OO
    from rgqa_b200.tasks.gqa_odin import mahalanobis_noise_scores
    sd, ig, inp = setup
    mu, P = syn.maha_stats(num_answers=60)
    ref_s, ref_l, ref_gf, ref_gb = OO.mahalanobis_with_noise(sd, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], mu, P, 1e-3)
    s, l, gf, gb = mahalanobis_noise_scores(ig, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], mu, P, 1e-3)
    assert _rel(gf.cpu(), ref_gf) <= 2e-4 and _rel(gb.cpu(), ref_gb) <= 2e-4
    assert torch.equal(l.cpu().long(), ref_l)
    torch.testing.assert_close(s.cpu(), ref_s, rtol=1e-3, atol=1e-5)


def test_full_depth_lxmert_input_gradients(cuda_device):
    """All 9 / 5 / 5 layers (five cross-modality layers in a row: gradients of both streams through the shared
    cross-attention module), B = 2, against autograd through the oracle."""
    from oracle import lxmert_oracle as O
    from rgqa_b200.lxrt.input_grad import LXMERTInputGrad
    sd = syn.lxmert_state_dict(variant="rich")
    ig = LXMERTInputGrad(sd, device=cuda_device)
    inp = syn.make_inputs(2, 20, seed=33)
    feats, boxes = inp["feat"].clone().requires_grad_(True), inp["pos"].clone().requires_grad_(True)
    x = O.lxmert_encode(sd, feats, boxes, inp["ids"], inp["seg"], inp["mask"])
    logit = O.gqa_head(sd, x)
    g = torch.Generator().manual_seed(6)
    wl = torch.randn(logit.shape, generator=g)
    (logit * wl).sum().backward()
    logits, _ = ig.forward(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"])
    assert (logits.cpu() - logit.detach()).abs().max().item() <= 1e-3
    dfeat, dpos = ig.backward(dlogits=wl)
    assert _rel(dfeat.cpu(), feats.grad) <= 5e-4, _rel(dfeat.cpu(), feats.grad)
    assert _rel(dpos.cpu(), boxes.grad) <= 5e-4, _rel(dpos.cpu(), boxes.grad)


def test_mixed_precision_input_gradients_and_odin(cuda_device):
    """gemm="bf16" (tensor-core matrix products, fp32 elsewhere): the input gradient stays within bf16-operand noise of autograd
    (5 % relative L2, >= 97 % sign agreement on the significant entries) and the ODIN scores within 1e-2 of the fp32 oracle."""
    from oracle import odin_oracle as OO
    from rgqa_b200.lxrt.input_grad import LXMERTInputGrad
    from rgqa_b200.tasks.gqa_odin import odin_scores
    sd = syn.lxmert_state_dict(variant="rich", **LAYERS)
    ig = LXMERTInputGrad(sd, device=cuda_device, gemm="bf16", **LAYERS)
    inp = syn.make_inputs(8, 20, seed=21)            # 160 / 288 token rows: above the tensor-core path's 64-row minimum
    ref_s, ref_l, ref_gf, ref_gb = OO.odin(sd, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], 1.0, 1e-3)
    s, l, gf, gb = odin_scores(ig, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], 1.0, 1e-3)
    assert _rel(gf.cpu(), ref_gf) <= 5e-2 and _rel(gb.cpu(), ref_gb) <= 5e-2, (_rel(gf.cpu(), ref_gf), _rel(gb.cpu(), ref_gb))
    big = ref_gf.abs() > 1e-2 * ref_gf.abs().max()
    assert (torch.sign(gf.cpu())[big] == torch.sign(ref_gf)[big]).float().mean().item() >= 0.97
    assert torch.equal(l.cpu().long(), ref_l)
    torch.testing.assert_close(s.cpu(), ref_s, rtol=1e-2, atol=1e-3)


def test_odin_and_mahalanobis_with_noise_against_the_reference_stack(cuda_device):
    """The two input-gradient scorers end to end against the UNMODIFIED reference: ``gqa_odin.GQA.ood_evaluate`` and
    ``gqa_maha.GQA.ood_evaluate`` (--noise 1e-3) run on the real ``GQAModel`` / ``GQAModel_maha`` with torch autograd through the
    reference LXMERT (tests/golden/make_golden_dropin.py), sentences tokenised by the product tokenizer, batches of the
    reference loader (the losses are batch means).  The perturbation moves the reference's own scores by up to 8.6e-5 (ODIN)
    and 0.084 (Mahalanobis); the CUDA path must land on the perturbed scores, not the plain ones
    (error at most a quarter of the shift)."""
    import gzip, json, os
    from rgqa_b200.lxrt.input_grad import LXMERTInputGrad
    from rgqa_b200.lxrt.tokenization import BertTokenizer
    from rgqa_b200.tasks.gqa_odin import mahalanobis_noise_scores, odin_scores
    gold = os.path.join(os.path.dirname(__file__), "golden")
    d = json.load(gzip.open(os.path.join(gold, "dropin_reference.json.gz"), "rt"))
    sd = syn.lxmert_state_dict(variant="rich", head_sigma=d["head_sigma"], **d["layers"])
    ig = LXMERTInputGrad(sd, device=cuda_device, **d["layers"])
    tok = BertTokenizer.from_pretrained(os.path.join(gold, "tokenizer_vocab.txt"), do_lower_case=True)
    ids, mask, seg = (torch.from_numpy(a).long() for a in tok.encode_batch(d["sents"], d["max_seq_length"]))
    inp = syn.make_inputs(len(d["qids"]), 20, seed=d["inputs_seed"])
    mu, P = syn.maha_stats()
    label2ans = [f"answer{j}" for j in range(d["num_answers"])]
    got = {"odin": ([], []), "maha_noise": ([], [])}
    for lo in range(0, len(d["qids"]), d["batch"]):
        sl = slice(lo, lo + d["batch"])
        s, l, _, _ = odin_scores(ig, inp["feat"][sl], inp["pos"][sl], ids[sl], seg[sl], mask[sl], 1.0, d["noise"])
        got["odin"][0].append(s.cpu()), got["odin"][1].append(l.cpu())
        s, l, _, _ = mahalanobis_noise_scores(ig, inp["feat"][sl], inp["pos"][sl], ids[sl], seg[sl], mask[sl], mu, P, d["noise"])
        got["maha_noise"][0].append(s.cpu()), got["maha_noise"][1].append(l.cpu())
    for name, plain, tol in (("odin", "msp", 4e-6), ("maha_noise", "maha", 2e-2)):
        s = torch.cat(got[name][0]).double().numpy()
        ans = [label2ans[int(i)] for i in torch.cat(got[name][1])]
        want, base = np.asarray(d["scores"][name]), np.asarray(d["scores"][plain])
        assert ans == d["answers"][name], name
        err = np.abs(s - want).max()
        shift = np.abs(want - base).max()
        assert err <= tol and shift >= 4 * err, (name, err, shift)


def test_sign_step_kernel(cuda_device):
    """x - noise * sign(grad) with sign(0) = +1 (gqa_odin.py:118-120: `gradient = torch.ge(grad, 0); gradient = (gradient.float() - 0.5) * 2`)."""
    from rgqa_b200.tasks.gqa_odin import _step_against_gradient_sign
    g = torch.Generator().manual_seed(3)
    for n in (7, 4096, 36 * 2048 * 3 + 5):
        x = torch.randn(n, generator=g).to(cuda_device)
        grad = torch.randn(n, generator=g).to(cuda_device)
        grad[::5] = 0.0                                    # ge(0, 0) is True: a zero gradient steps DOWN
        grad[1::10] = -0.0
        for noise in (1e-3, 1e-4, 0.0):
            y = _step_against_gradient_sign(x, grad, noise)
            gradient = (torch.ge(grad, 0).float() - 0.5) * 2   # the reference's two statements, literally
            want = torch.add(x, gradient, alpha=-noise)
            torch.testing.assert_close(y, want, rtol=0, atol=1e-7)


@pytest.mark.parametrize("kind", ["odin", "maha"])
def test_graph_replay_of_the_input_gradient_scorers_equals_the_eager_pass(setup, kind):
    """`GraphedInputGradScorer`: forward + input-gradient backward + sign step + forward captured once per batch shape and replayed;
    four different batches through the same graph must give the eager scorer's scores (same kernels; the scorers have no
    atomics on this path, so equality is exact) and labels."""
    from rgqa_b200.tasks.gqa_odin import GraphedInputGradScorer, mahalanobis_noise_scores, odin_scores
    sd, ig, _ = setup
    mu, P = syn.maha_stats(num_answers=60)
    gs = GraphedInputGradScorer(ig, kind, temperature=2.0, noise=1e-3, sample_mean=mu, precision=P)
    for step in range(5):
        inp = syn.make_inputs(3, 20, seed=100 + step)
        a = [inp[k] for k in ("feat", "pos", "ids", "seg", "mask")]
        if kind == "odin":
            ws, wl, _, _ = odin_scores(ig, *a, 2.0, 1e-3)
        else:
            ws, wl, _, _ = mahalanobis_noise_scores(ig, *a, mu, P, 1e-3)
        ws, wl = ws.clone(), wl.clone()
        s, l = gs(*a)
        torch.cuda.synchronize()
        assert torch.equal(l, wl) and torch.equal(s, ws), (step, (s - ws).abs().max().item())
    assert gs.captures == 1
