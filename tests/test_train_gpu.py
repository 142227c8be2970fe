"""SURVEY §8 row a22 / config 5 on the GPU: the weight-gradient kernels against torch, and three optimisation steps of the
trainer against the UNMODIFIED reference modules + reference BertAdam (tests/golden/train_steps.npz, dropout off)."""
import os

import numpy as np
import pytest
import torch

from rgqa_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_wgrad_kernels_match_torch(cuda_device):
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(3)
    for M, N, K in ((640, 768, 768), (1152, 2304, 768), (6, 1842, 1536), (216, 768, 4), (77, 130, 96)):
        dy, x = torch.randn(M, N, generator=g), torch.randn(M, K, generator=g)
        dw0, db0 = torch.randn(N, K, generator=g), torch.randn(N, generator=g)          # kernels accumulate
        dw, db = dw0.clone().to(cuda_device), db0.clone().to(cuda_device)
        ops.linear_wgrad_(dw, db, dy.to(cuda_device), x.to(cuda_device))
        ref_w = dw0.double() + dy.double().t() @ x.double()
        ref_b = db0.double() + dy.double().sum(0)
        assert (dw.cpu().double() - ref_w).abs().max() <= 1e-4 * max(1.0, float(ref_w.abs().max()))
        assert (db.cpu().double() - ref_b).abs().max() <= 1e-4 * max(1.0, float(ref_b.abs().max()))
    # a strided dy (one third of a fused Q|K|V gradient)
    dqkv, x = torch.randn(300, 2304, generator=g).to(cuda_device), torch.randn(300, 768, generator=g).to(cuda_device)
    dw, db = torch.zeros(768, 768, device=cuda_device), torch.zeros(768, device=cuda_device)
    ops.linear_wgrad_(dw, db, dqkv[:, 768:1536], x)
    torch.testing.assert_close(dw, dqkv[:, 768:1536].t() @ x, rtol=1e-4, atol=1e-3)
    for H in (768, 1536):
        xx = torch.randn(333, H, generator=g, requires_grad=True)
        gamma, beta = torch.randn(H, generator=g, requires_grad=True), torch.randn(H, generator=g, requires_grad=True)
        dy = torch.randn(333, H, generator=g)
        y = torch.nn.functional.layer_norm(xx, (H,), gamma, beta, eps=1e-12)
        (y * dy).sum().backward()
        dg, dbt = torch.zeros(H, device=cuda_device), torch.zeros(H, device=cuda_device)
        ops.layernorm_param_grad_(dg, dbt, dy.to(cuda_device), xx.detach().to(cuda_device))
        torch.testing.assert_close(dg.cpu(), gamma.grad, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(dbt.cpu(), beta.grad, rtol=1e-4, atol=1e-4)
        dx = ops.layernorm_bwd(dy.to(cuda_device), xx.detach().to(cuda_device), gamma.detach().to(cuda_device))
        torch.testing.assert_close(dx.cpu(), xx.grad, rtol=1e-4, atol=1e-5)
    table = torch.zeros(50, 768, device=cuda_device)
    idx = torch.randint(0, 50, (400,), generator=g).to(torch.int32)
    gg = torch.randn(400, 768, generator=g)
    ops.embedding_scatter_(table, idx.to(cuda_device), gg.to(cuda_device))
    ref = torch.zeros(50, 768).index_add_(0, idx.long(), gg)
    torch.testing.assert_close(table.cpu(), ref, rtol=1e-5, atol=1e-5)


def test_three_training_steps_match_reference_trajectory(cuda_device):
    """gqa_conf.py:148-198 with dropout off: loss and pre-clip gradient norm of each step, first-step gradients and the
    parameters after step 3 of twelve watched tensors, against the reference modules + reference BertAdam."""
    from rgqa_b200.lxrt.trainer import LXMERTTrainer
    z = np.load(os.path.join(GOLD, "train_steps.npz"))
    layers = dict(l_layers=2, x_layers=1, r_layers=1)
    sd = syn.lxmert_state_dict(variant="rich", **layers)
    tr = LXMERTTrainer(sd, device=cuda_device, lr=1e-4, warmup=0.1, t_total=20, **layers)
    watch = sorted(k[len("final:"):] for k in z.files if k.startswith("final:"))
    assert len(watch) == 12

    def cut(t):
        return t[:4] if t.numel() > 20000 else t

    for step in range(3):
        inp = syn.make_inputs(6, 20, seed=900 + step)
        loss, norm = tr.train_step(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], torch.from_numpy(z[f"target{step}"]))
        assert loss == pytest.approx(float(z[f"loss{step}"]), rel=2e-5), (step, loss)
        assert norm == pytest.approx(float(z[f"norm{step}"]), rel=2e-4), (step, norm)
        if step == 0:
            for n in watch:
                ref = torch.from_numpy(z["grad0:" + n])
                got = cut(tr.g[n]).cpu()
                if float(ref.norm()) == 0.0:      # token types are all 0 = padding_idx: the reference leaves that table's gradient at zero
                    assert float(got.norm()) == 0.0, n
                    continue
                assert float((got - ref).norm() / ref.norm()) <= 5e-4, (n, float((got - ref).norm() / ref.norm()))
            rows = tr.g["lxrt_encoder.model.bert.embeddings.word_embeddings.weight"][[101, 102, 0]].cpu()
            ref = torch.from_numpy(z["grad0_word_rows"])
            assert float((rows - ref).norm() / ref.norm()) <= 5e-4
    for n in watch:
        ref = torch.from_numpy(z["final:" + n])
        got = cut(tr.p[n]).cpu()
        start = cut(sd[n])
        moved = float((ref - start).norm())
        assert moved > 0
        assert float((got - ref).norm()) <= 2e-3 * moved + 1e-7, (n, float((got - ref).norm()), moved)   # the UPDATE agrees to 0.2 %


def test_mixed_precision_training_step_tracks_the_fp32_trajectory(cuda_device):
    """gemm="bf16": forward / dgrad / wgrad matrix products on the tcgen05 tensor cores with bf16 operands (fp32 master weights,
    activations, gradients, optimiser).  Not the reference's arithmetic: the same three steps must stay within bf16-operand noise
    of the reference trajectory — loss 2e-3, gradient norm 2 %, watched gradients 5 % (relative L2), updates 10 %."""
    from rgqa_b200.lxrt.trainer import LXMERTTrainer
    z = np.load(os.path.join(GOLD, "train_steps.npz"))
    layers = dict(l_layers=2, x_layers=1, r_layers=1)
    sd = syn.lxmert_state_dict(variant="rich", **layers)
    tr = LXMERTTrainer(sd, device=cuda_device, lr=1e-4, warmup=0.1, t_total=20, gemm="bf16", **layers)
    watch = sorted(k[len("final:"):] for k in z.files if k.startswith("final:"))
    cut = lambda t: t[:4] if t.numel() > 20000 else t
    for step in range(3):
        inp = syn.make_inputs(6, 20, seed=900 + step)
        loss, norm = tr.train_step(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], torch.from_numpy(z[f"target{step}"]))
        assert loss == pytest.approx(float(z[f"loss{step}"]), rel=2e-3), (step, loss)
        assert norm == pytest.approx(float(z[f"norm{step}"]), rel=2e-2), (step, norm)
        if step == 0:
            for n in watch:
                ref = torch.from_numpy(z["grad0:" + n])
                if float(ref.norm()) == 0.0:
                    continue
                got = cut(tr.g[n]).cpu()
                assert float((got - ref).norm() / ref.norm()) <= 5e-2, (n, float((got - ref).norm() / ref.norm()))
    for n in watch:
        ref, got, start = torch.from_numpy(z["final:" + n]), cut(tr.p[n]).cpu(), cut(sd[n])
        moved = float((ref - start).norm())
        assert float((got - ref).norm()) <= 0.1 * moved + 1e-7, (n, float((got - ref).norm()), moved)


def test_dropout_kernels_against_torch_with_the_same_mask(cuda_device):
    """Counter-based dropout: keep rate, scaling, determinism; attention with dropout on the probabilities and its backward against
    torch autograd fed the SAME mask (recovered by running the elementwise kernel on ones with the attention kernel's seed)."""
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.randn(1 << 20, generator=g).to(cuda_device)
    r = torch.randn(1 << 20, generator=g).to(cuda_device)
    y = ops.dropout_add(x, None, 0.1, seed=77)
    kept = (y != 0)
    assert abs(float(kept.float().mean()) - 0.9) < 2e-3
    torch.testing.assert_close(y[kept], (x / 0.9)[kept], rtol=1e-6, atol=0)
    assert torch.equal(y, ops.dropout_add(x, None, 0.1, seed=77)) and not torch.equal(y, ops.dropout_add(x, None, 0.1, seed=78))
    torch.testing.assert_close(ops.dropout_add(x, r, 0.1, seed=77), y + r)
    B, Lq, Lk, heads, H, p, seed = 3, 20, 36, 12, 768, 0.1, 4242
    qkv_q = torch.randn(B * Lq, 3 * H, generator=g).to(cuda_device)
    qkv_kv = torch.randn(B * Lk, 3 * H, generator=g).to(cuda_device)
    lens = torch.randint(1, Lk + 1, (B,), generator=g)
    mask = (torch.arange(Lk)[None, :] < lens[:, None]).to(torch.int32).to(cuda_device)
    out = ops.attention_dropout(qkv_q, qkv_kv, B, Lq, Lk, key_mask=mask, heads=heads, p=p, seed=seed)
    keep = ops.dropout_add(torch.ones(B * heads * Lq * Lk, device=cuda_device), None, p, seed).view(B, heads, Lq, Lk)   # 0 or 1 / (1 - p)
    q = qkv_q[:, :H].clone().requires_grad_(True)
    k = qkv_kv[:, H:2 * H].clone().requires_grad_(True)
    v = qkv_kv[:, 2 * H:].clone().requires_grad_(True)
    ql, kl, vl = (t.view(B, -1, heads, 64).permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(ql, kl.transpose(-1, -2)) / 8.0 + ((1.0 - mask.float()) * -10000.0)[:, None, None, :]
    ref = torch.matmul(torch.softmax(s, dim=-1) * keep, vl).permute(0, 2, 1, 3).reshape(B * Lq, H)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
    dctx = torch.randn(B * Lq, H, generator=g).to(cuda_device)
    (ref * dctx).sum().backward()
    dqkv_q, dqkv_kv = torch.zeros_like(qkv_q), torch.zeros_like(qkv_kv)
    ops.attention_bwd(qkv_q, qkv_kv, dctx, dqkv_q, dqkv_kv, B, Lq, Lk, key_mask=mask, heads=heads, p=p, seed=seed)
    torch.testing.assert_close(dqkv_q[:, :H], q.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(dqkv_kv[:, H:2 * H], k.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(dqkv_kv[:, 2 * H:], v.grad, rtol=1e-3, atol=1e-5)


def test_training_step_with_dropout_backward_matches_its_forward(cuda_device):
    """dropout = 0.1 (the reference's training configuration; masks are this package's, not torch's): with the masks frozen by the
    seed, the analytic gradient must be the derivative of the loss the forward computes — checked as a directional derivative along
    a random parameter direction with central differences."""
    from rgqa_b200 import _lib as L, ops
    from rgqa_b200.lxrt.trainer import LXMERTTrainer
    layers = dict(l_layers=1, x_layers=1, r_layers=1)
    sd = syn.lxmert_state_dict(variant="rich", **layers)
    tr = LXMERTTrainer(sd, device=cuda_device, dropout=0.1, seed=5, **layers)
    inp = syn.make_inputs(4, 20, seed=17)
    g = torch.Generator().manual_seed(3)
    target = torch.zeros(4, 1842)
    target[torch.arange(4), torch.randint(0, 1842, (4,), generator=g)] = 1.0

    def loss_only():
        logits, _ = tr.forward(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"])
        loss = torch.zeros(1, dtype=torch.float64, device=tr.dev)
        dlog = torch.empty_like(logits)
        t = target.to(tr.dev)
        L.check(L.load().rgqa_bce_loss_grad(L.current_stream_ptr(), L.ptr(logits), L.ptr(t), 4, 1842, L.ptr(dlog), L.ptr(loss)), "bce")
        return float(loss), dlog

    l0, dlog = loss_only()
    l0b, _ = loss_only()
    assert l0 == l0b                                           # same step -> same masks -> same loss
    tr.flat_g.zero_()
    tr.backward(dlog)
    names = ["lxrt_encoder.model.bert.encoder.x_layers.0.visual_attention.att.query.weight",
             "lxrt_encoder.model.bert.encoder.layer.0.output.dense.weight",
             "lxrt_encoder.model.bert.encoder.r_layers.0.attention.output.LayerNorm.weight", "logit_fc.0.weight"]
    for n in names:
        d = torch.randn(tr.p[n].shape, generator=g).to(tr.dev)
        d = d / d.norm()
        analytic = float((tr.g[n] * d).sum())
        eps = 2e-2
        tr.p[n].add_(eps * d); tr._derived.clear()
        lp, _ = loss_only()
        tr.p[n].sub_(2 * eps * d); tr._derived.clear()
        lm, _ = loss_only()
        tr.p[n].add_(eps * d); tr._derived.clear()
        numeric = (lp - lm) / (2 * eps)
        assert numeric == pytest.approx(analytic, rel=5e-2, abs=5e-3), (n, numeric, analytic)
    # and the no-dropout trainer differs (the masks really are applied)
    tr0 = LXMERTTrainer(sd, device=cuda_device, **layers)
    lg0, _ = tr0.forward(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"])
    lg1, _ = tr.forward(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"])
    assert (lg0 - lg1).abs().max().item() > 1e-3


def test_rp_finetune_loop_kernels_follow_the_reference_train_loop(cuda_device):
    """The trainer's own pieces — random-pairing batch (``tasks/rp.py``), ``BCEWithLogitsLoss x A`` + d/dlogit kernel, global-norm
    clip 5.0 and the BertAdam kernel with the ``warmup_linear`` schedule over ``t_total = batches x epochs`` — driven around the tiny
    stand-in model of tests/golden/tiny_vqa.py (torch autograd supplies its gradients: test scaffolding), against six steps of the
    UNMODIFIED reference ``GQA.train`` loop (src/tasks/gqa_conf.py:140-245, tests/golden/make_golden_rp_train.py)."""
    import json, random, sys
    from rgqa_b200.lxrt.optimization import BertAdam, bce_loss_and_grad
    from rgqa_b200.tasks.rp import build_rp_batch
    gold = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, gold)
    from tiny_vqa import TinyVQA
    z = np.load(os.path.join(gold, "rp_train_reference.npz"))
    with open(os.path.join(gold, "rp_train_reference.json")) as f:
        meta = json.load(f)
    sd = {k[len("init."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init.")}
    C, D = sd["head.weight"].shape
    model = TinyVQA(sd["fc_feat.weight"].shape[1], D, C)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_device).train()
    assert [k for k, _ in model.named_parameters()] == meta["param_order"]
    params = [p.data for p in model.parameters()]
    qids, sents, img = meta["qids"], meta["sents"], meta["img_of"]
    feats, boxes, target = (torch.from_numpy(z[k]) for k in ("feat", "box", "target"))
    bs = meta["batch"]
    n_batches = (len(qids) + bs - 1) // bs
    opt = BertAdam(params, lr=meta["lr"], warmup=0.1, t_total=n_batches * meta["epochs"])
    random.seed(meta["seed"])
    step = 0
    for _ in range(meta["epochs"]):
        for lo in range(0, len(qids), bs):
            sl = slice(lo, lo + bs)
            _, f2, b2, s2, t2 = build_rp_batch(qids[sl], feats[sl], boxes[sl], sents[sl], target[sl][:, :-1], img.__getitem__)
            assert s2 == meta["steps"][step]["sent"]
            model.zero_grad(set_to_none=True)
            logit = model(f2.to(cuda_device), b2.to(cuda_device), s2)
            loss, dlogit = bce_loss_and_grad(logit.detach().float().contiguous(), t2.to(cuda_device).float().contiguous())
            logit.backward(dlogit)
            opt.step([p.grad.contiguous() for p in model.parameters()], clip_norm=5.0)
            torch.cuda.synchronize()
            want_loss = meta["steps"][step]["loss_times_A"]
            assert abs(float(loss.item()) - want_loss) <= 2e-5 * abs(want_loss), (step, float(loss.item()), want_loss)
            got = torch.cat([p.reshape(-1) for p in params]).cpu().numpy()
            np.testing.assert_allclose(got, z["params_after"][step], rtol=2e-4, atol=1e-5, err_msg=f"step {step}")
            step += 1
    assert step == len(meta["steps"]) == 6
    # the six updates moved the parameters by far more than the tolerance (the comparison is not vacuous)
    init = torch.cat([sd[k].reshape(-1) for k in meta["param_order"]]).numpy()
    assert np.abs(z["params_after"][-1] - init).max() > 5e-3


@pytest.mark.parametrize("gemm", ["fp32", "bf16"])
def test_cuda_graph_replay_of_the_gradient_pass_matches_eager_steps(cuda_device, gemm):
    """`use_graph=True` captures zero-grad + forward + loss + backward once per input shape and replays it (the ~1 000 launches of
    a step otherwise come from Python); the optimiser and the schedule stay outside.  Five steps — eager, capture, three
    replays — must leave the same losses, gradient norms and parameters as five eager steps (up to the run-to-run order of the
    weight-gradient atomics)."""
    from rgqa_b200.lxrt.trainer import LXMERTTrainer
    layers = dict(l_layers=2, x_layers=1, r_layers=1)
    sd = syn.lxmert_state_dict(variant="rich", **layers)
    runs = []
    for use_graph in (False, True):
        tr = LXMERTTrainer(sd, device=cuda_device, lr=1e-4, warmup=0.1, t_total=20, gemm=gemm, use_graph=use_graph, **layers)
        log = []
        for step in range(5):
            inp = syn.make_inputs(8, 20, seed=300 + step)
            g = torch.Generator().manual_seed(step)
            target = torch.zeros(8, 1842)
            target[torch.arange(8), torch.randint(0, 1842, (8,), generator=g)] = 1.0
            log.append(tr.train_step(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], target))
        runs.append((log, tr.flat_p.clone()))
    # the weight-gradient kernels accumulate with float atomics (order varies run to run, graph or not): equality to ~1e-9
    # (bf16 GEMMs: an fp32 weight that differs in its last bit can round to the neighbouring bf16 value, which amplifies the atomics'
    # run-to-run noise along the trajectory — two EAGER runs differ by as much: 5e-6 on the loss after four updates)
    # BertAdam normalises every element's update to ~lr (1e-4 here), so an element whose gradient is pure rounding noise can move
    # by a different +-lr in the two runs: element-wise 2 lr, and the difference as a whole must stay at the noise level
    rl, rn, ap = (1e-7, 1e-6, 2e-7) if gemm == "fp32" else (5e-5, 5e-4, 2e-4)
    for (l0, n0), (l1, n1) in zip(runs[0][0], runs[1][0]):
        assert l1 == pytest.approx(l0, rel=rl) and n1 == pytest.approx(n0, rel=rn), (runs[0][0], runs[1][0])
    assert runs[0][0][0][0] == runs[1][0][0][0] and runs[0][0][1][0] == pytest.approx(runs[1][0][1][0], rel=1e-7)   # first steps: tight
    torch.testing.assert_close(runs[1][1], runs[0][1], rtol=0, atol=ap)
    # (bf16: the weight-gradient GEMMs split K over CTAs and add atomically — more run-to-run summation-order noise than the fp32 kernels)
    assert float((runs[1][1] - runs[0][1]).norm() / runs[0][1].norm()) <= (1e-5 if gemm == "fp32" else 3e-5)


@pytest.mark.parametrize("R,C,ld", [(2560, 768, 768), (4608, 768, 2304), (130, 70, 70), (7, 3072, 3072), (64, 64, 200)])
def test_transpose_cast_kernel(cuda_device, R, C, ld):
    """bf16 [C, R] = transpose of an fp32 [R, C] view (row stride ld): the weight-gradient operand re-layout in one pass."""
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(R + C)
    base = torch.randn(R, ld, generator=g).to(cuda_device)
    x = base[:, ld - C:] if ld > C else base
    y = ops.transpose_cast_bf16(x)
    assert y.shape == (C, R) and torch.equal(y, x.t().contiguous().to(torch.bfloat16))


@pytest.mark.parametrize("B,Lq,Lk,masked", [(3, 20, 20, True), (2, 36, 36, False), (2, 20, 36, False), (2, 36, 20, True),
                                            (1, 64, 64, True), (2, 17, 7, True), (1, 1, 5, False), (2, 56, 56, True)])
def test_attention_backward_kernel_against_autograd(cuda_device, B, Lq, Lk, masked):
    """dQ / dK / dV of softmax(QK^T/8 + mask) V (src/lxrt/modeling.py:330-346) for every (Lq, Lk) of the path, ragged lengths
    (not multiples of the 4 x 4 register tile) and the 64-row maximum, against torch autograd in fp64."""
    from rgqa_b200 import ops
    heads, H = 12, 768
    g = torch.Generator().manual_seed(B * 1000 + Lq * 10 + Lk)
    qkv_q = torch.randn(B * Lq, 3 * H, generator=g).to(cuda_device)
    qkv_kv = torch.randn(B * Lk, 3 * H, generator=g).to(cuda_device)
    mask = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), generator=g)
        mask = (torch.arange(Lk)[None, :] < lens[:, None]).to(torch.int32).to(cuda_device)
    dctx = torch.randn(B * Lq, H, generator=g).to(cuda_device)
    q = qkv_q[:, :H].double().clone().requires_grad_(True)
    k = qkv_kv[:, H:2 * H].double().clone().requires_grad_(True)
    v = qkv_kv[:, 2 * H:].double().clone().requires_grad_(True)
    ql, kl, vl = (t.view(B, -1, heads, 64).permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(ql, kl.transpose(-1, -2)) / 8.0
    if mask is not None:
        s = s + ((1.0 - mask.double()) * -10000.0)[:, None, None, :]
    ref = torch.matmul(torch.softmax(s, dim=-1), vl).permute(0, 2, 1, 3).reshape(B * Lq, H)
    (ref * dctx.double()).sum().backward()
    dqkv_q = torch.full_like(qkv_q, float("nan"))
    dqkv_kv = torch.full_like(qkv_kv, float("nan"))
    ops.attention_bwd(qkv_q, qkv_kv, dctx, dqkv_q, dqkv_kv, B, Lq, Lk, key_mask=mask, heads=heads)
    torch.cuda.synchronize()
    torch.testing.assert_close(dqkv_q[:, :H].double(), q.grad, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(dqkv_kv[:, H:2 * H].double(), k.grad, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(dqkv_kv[:, 2 * H:].double(), v.grad, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("M,H,with_add", [(2560, 768, False), (4608, 768, True), (37, 768, True), (128, 1536, False)])
def test_layernorm_backward_with_parameter_gradients_in_one_pass(cuda_device, M, H, with_add):
    """rgqa_layernorm_bwd_params == rgqa_layernorm_bwd + rgqa_ln_param_grad (BertLayerNorm backward, src/lxrt/modeling.py:258-270)
    and both against torch autograd in fp64."""
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(M + H)
    x = (torch.randn(M, H, generator=g) * 1.7 + 0.3).to(cuda_device)
    dy = torch.randn(M, H, generator=g).to(cuda_device)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).to(cuda_device)
    add = torch.randn(M, H, generator=g).to(cuda_device) if with_add else None
    dg0 = torch.randn(H, generator=g).to(cuda_device)          # the kernels ACCUMULATE into the parameter gradients
    db0 = torch.randn(H, generator=g).to(cuda_device)
    dg, db = dg0.clone(), db0.clone()
    dx = ops.layernorm_bwd_params_(dg, db, dy, x, gamma, add=add, scale=0.5)
    dg2, db2 = dg0.clone(), db0.clone()
    ops.layernorm_param_grad_(dg2, db2, dy, x, 1e-12, 0.5)
    dx2 = ops.layernorm_bwd(dy, x, gamma, add=add, scale=0.5)
    torch.cuda.synchronize()
    torch.testing.assert_close(dx, dx2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dg, dg2, rtol=1e-4, atol=2e-3)   # summation order of M rows (float atomics)
    torch.testing.assert_close(db, db2, rtol=1e-4, atol=2e-3)
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = torch.zeros(H, dtype=torch.float64, device=cuda_device, requires_grad=True)
    y = torch.nn.functional.layer_norm(xd, (H,), gd, bd, eps=1e-12)
    (y * dy.double() * 0.5).sum().backward()
    want_dx = xd.grad + (add.double() if with_add else 0.0)
    torch.testing.assert_close(dx.double(), want_dx, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close((dg - dg0).double(), gd.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close((db - db0).double(), bd.grad, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("R,C,ld", [(2560, 768, 768), (4608, 768, 2304), (130, 70, 70), (64, 3072, 3072)])
def test_transpose_cast_kernel_with_column_sums(cuda_device, R, C, ld):
    """The bias gradient rides on the re-layout of dy: colsum += sum over rows, same bf16 transpose."""
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(R + C + 1)
    base = torch.randn(R, ld, generator=g).to(cuda_device)
    x = base[:, ld - C:] if ld > C else base
    acc0 = torch.randn(C, generator=g).to(cuda_device)
    acc = acc0.clone()
    y = ops.transpose_cast_bf16(x, colsum_into=acc)
    torch.cuda.synchronize()
    assert torch.equal(y, x.t().contiguous().to(torch.bfloat16))
    torch.testing.assert_close((acc - acc0).double(), x.double().sum(0), rtol=1e-5, atol=2e-4)


@pytest.mark.parametrize("M,N,K", [(768, 768, 4608), (3072, 768, 2560), (768, 3072, 4608), (768, 768, 2560), (1842, 1536, 128),
                                   (1536, 768, 128), (2304, 768, 4608), (130, 70, 64)])
def test_accumulating_weight_gradient_gemm(cuda_device, M, N, K):
    """out += A · Wᵀ with the K range split over CTAs and atomic accumulation (TMA reduce-add in the CTA-pair kernel, red.global in
    the single-CTA one): the weight-gradient product dW += dYᵀ · X of `gqa_conf.py:196`'s backward at the RP-finetune shapes."""
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_device)
    w = torch.randn(N, K, generator=g).to(torch.bfloat16).to(cuda_device)
    out0 = torch.randn(M, N, generator=g).to(cuda_device)
    out = out0.clone()
    ops.linear_bf16_accumulate_(out, a, w)
    ops.linear_bf16_accumulate_(out, a, w)               # accumulates: twice the product
    torch.cuda.synchronize()
    ref = out0.double() + 2.0 * (a.double() @ w.double().t())
    torch.testing.assert_close(out.double(), ref, rtol=1e-5, atol=2e-3 * (K / 128) ** 0.5)   # fp32 accumulation order only


@pytest.mark.parametrize("M,N,K", [(768, 768, 4608), (3072, 768, 2560), (768, 3072, 4608), (2304, 768, 128), (1840, 1536, 128),
                                   (768, 2048, 4608), (768, 64, 2560)])
def test_weight_gradient_gemm_from_transposed_operands(cuda_device, M, N, K):
    """out += Atᵀ · Bt with At [K, M], Bt [K, N] row-major bf16 read MN-major by the tensor cores (dW += dYᵀ · X without any
    transposing copy), strided operands (dY as a column slice of the stacked Q|K|V gradient), twice to check the accumulation."""
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    wide = torch.randn(K, M + 256, generator=g).to(torch.bfloat16).to(cuda_device)
    at = wide[:, 128:128 + M]                                            # row stride M + 256, 256-byte aligned start
    bt = torch.randn(K, N, generator=g).to(torch.bfloat16).to(cuda_device)
    out0 = torch.randn(M, N, generator=g).to(cuda_device)
    out = out0.clone()
    ops.linear_bf16_accumulate_tn_(out, at, bt)
    ops.linear_bf16_accumulate_tn_(out, at, bt)
    torch.cuda.synchronize()
    ref = out0.double() + 2.0 * (at.double().t() @ bt.double())
    torch.testing.assert_close(out.double(), ref, rtol=1e-5, atol=2e-3 * (K / 128) ** 0.5)


@pytest.mark.parametrize("R,C,ld", [(2560, 768, 768), (4608, 768, 2304), (130, 72, 72), (128, 1844, 1844)])
def test_cast_with_column_sums(cuda_device, R, C, ld):
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(R + C)
    base = torch.randn(R, ld, generator=g).to(cuda_device)
    x = base[:, ld - C:] if ld > C else base
    acc0 = torch.randn(C, generator=g).to(cuda_device)
    acc = acc0.clone()
    y = ops.cast_bf16_colsum(x, colsum_into=acc)
    y2 = ops.cast_bf16_colsum(x)
    wide = torch.zeros(R, C + 16, dtype=torch.bfloat16, device=cuda_device)
    ops.cast_bf16_colsum(x, out=wide[:, 8:8 + C])                       # into a column block of a wider buffer
    torch.cuda.synchronize()
    assert torch.equal(y, x.to(torch.bfloat16)) and torch.equal(y2, y) and torch.equal(wide[:, 8:8 + C], y)
    assert not wide[:, :8].any() and not wide[:, 8 + C:].any()
    torch.testing.assert_close((acc - acc0).double(), x.double().sum(0), rtol=1e-5, atol=2e-4)


@pytest.mark.parametrize("B,Lq,Lk,masked", [(3, 20, 20, True), (2, 36, 36, False), (2, 20, 36, False), (2, 36, 20, True),
                                            (1, 64, 64, True), (2, 17, 7, True), (1, 1, 5, False), (2, 56, 56, True), (128, 36, 36, False)])
def test_tensor_core_attention_backward(cuda_device, B, Lq, Lk, masked):
    """`rgqa_attention_bwd_bf16` (bf16 mma.sync products, fp32 softmax / accumulation) against torch autograd in fp64 on the bf16
    values of its operands, and against the fp32 kernel: every (Lq, Lk) of the path, ragged lengths, the 64-row maximum."""
    from rgqa_b200 import ops
    heads, H = 12, 768
    g = torch.Generator().manual_seed(B * 1000 + Lq * 10 + Lk + 1)
    qkv_q = torch.randn(B * Lq, 3 * H, generator=g).to(torch.bfloat16).float().to(cuda_device)     # bf16-representable: the cast is exact
    qkv_kv = torch.randn(B * Lk, 3 * H, generator=g).to(torch.bfloat16).float().to(cuda_device)
    mask = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), generator=g)
        mask = (torch.arange(Lk)[None, :] < lens[:, None]).to(torch.int32).to(cuda_device)
    dctx = torch.randn(B * Lq, H, generator=g).to(torch.bfloat16).float().to(cuda_device)
    q = qkv_q[:, :H].double().clone().requires_grad_(True)
    k = qkv_kv[:, H:2 * H].double().clone().requires_grad_(True)
    v = qkv_kv[:, 2 * H:].double().clone().requires_grad_(True)
    ql, kl, vl = (t.view(B, -1, heads, 64).permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(ql, kl.transpose(-1, -2)) / 8.0
    if mask is not None:
        s = s + ((1.0 - mask.double()) * -10000.0)[:, None, None, :]
    ref = torch.matmul(torch.softmax(s, dim=-1), vl).permute(0, 2, 1, 3).reshape(B * Lq, H)
    (ref * dctx.double()).sum().backward()
    dq_kv = {}
    for bf16 in (True, False):
        dqkv_q = torch.full_like(qkv_q, float("nan"))
        dqkv_kv = torch.full_like(qkv_kv, float("nan"))
        ops.attention_bwd(qkv_q, qkv_kv, dctx, dqkv_q, dqkv_kv, B, Lq, Lk, key_mask=mask, heads=heads, bf16=bf16)
        torch.cuda.synchronize()
        dq_kv[bf16] = (dqkv_q[:, :H].double(), dqkv_kv[:, H:2 * H].double(), dqkv_kv[:, 2 * H:].double())
    for got, want in zip(dq_kv[True], (q.grad, k.grad, v.grad)):
        assert torch.isfinite(got).all()
        # P and dS are rounded to bf16 before the second products: 2^-9 relative per term, random signs
        err = (got - want).abs().max().item()
        assert err <= 2e-2 * max(1.0, want.abs().max().item()), (err, want.abs().max().item())
        assert float((got - want).norm() / want.norm()) <= 6e-3
    for got, want in zip(dq_kv[True], dq_kv[False]):
        assert float((got - want).norm() / want.norm()) <= 6e-3


def test_mixed_precision_step_at_64_row_multiples_tracks_the_fp32_step(cuda_device):
    """B = 16 (320 language rows, 576 visual rows: multiples of 64) switches the mixed-precision step onto its tensor-core forms
    everywhere — bf16 attention forward and backward, LayerNorm / GELU kernels that leave the next GEMM's bf16 operand, weight
    gradients from transposed operands with K split over CTA pairs, one cast + bias-gradient pass per dy.  Against the fp32 step of the
    same trainer class (itself pinned to the reference trajectory above): loss 2e-3, gradient norm 2 %, every parameter tensor's
    gradient 6 % in relative L2 (biases and LayerNorm parameters included), flat gradient cosine >= 0.999."""
    from rgqa_b200.lxrt.trainer import LXMERTTrainer
    layers = dict(l_layers=2, x_layers=2, r_layers=1)
    sd = syn.lxmert_state_dict(variant="rich", **layers)
    inp = syn.make_inputs(16, 20, seed=31)
    g = torch.Generator().manual_seed(7)
    target = torch.zeros(16, 1842)
    target[torch.arange(16), torch.randint(0, 1842, (16,), generator=g)] = 1.0
    out = {}
    for gemm in ("fp32", "bf16"):
        tr = LXMERTTrainer(sd, device=cuda_device, lr=1e-4, warmup=0.1, t_total=20, gemm=gemm, **layers)
        loss = tr._grad_pass(inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"], target)
        torch.cuda.synchronize()
        out[gemm] = (float(loss), tr.flat_g.clone(), {n: t.clone() for n, t in tr.g.items()})
    (l32, f32, g32), (l16, f16, g16) = out["fp32"], out["bf16"]
    assert l16 == pytest.approx(l32, rel=2e-3)
    assert float(f16.norm()) == pytest.approx(float(f32.norm()), rel=2e-2)
    assert float((f16 * f32).sum() / (f16.norm() * f32.norm())) >= 0.999
    # (key biases have a mathematically zero gradient — softmax ignores a shift of all scores of a row — so theirs is rounding noise
    # in both runs; tensors below 1e-4 of the global norm are compared on the absolute scale instead)
    floor = 1e-4 * float(f32.norm())
    rel = sorted(((float((g16[n] - g32[n]).norm() / g32[n].norm()), n) for n in g32 if float(g32[n].norm()) > floor), reverse=True)
    assert len(rel) >= 0.7 * len(g32) and rel[0][0] <= 6e-2, rel[:5]
    small = max((float((g16[n] - g32[n]).norm()) for n in g32 if float(g32[n].norm()) <= floor), default=0.0)
    assert small <= floor, small
