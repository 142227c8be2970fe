"""BASELINE north_star item 3 on the GPU: the answer head's last GEMM with the MSP / energy / logsumexp reductions in its
epilogue, the tensor-core (split-bf16) Mahalanobis with the arg-max in its epilogue, and `rgqa_forward_scored` that strings
them behind the forward — against the stand-alone scorer kernels, the fp32 CUDA-core Mahalanobis and the CPU oracle
(src/tasks/gqa_conf.py:326, gqa_energy.py:131,185,204-206, gqa_maha.py:228-236,270)."""
import sys
import types

import numpy as np
import pytest
import torch

from rgqa_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _reduce_logits(a, w, bias, flags, want_logits):
    from rgqa_b200 import _lib as L
    lib = L.load()
    M, K = a.shape
    N = w.shape[0]
    work = torch.empty(int(lib.rgqa_reduce_work_bytes(M, N)), dtype=torch.uint8, device=a.device)
    scores = torch.full((4, M), float("nan"), dtype=torch.float32, device=a.device)
    labels = torch.full((M,), -1, dtype=torch.int32, device=a.device)
    logits = torch.empty((M, N), dtype=torch.float32, device=a.device) if want_logits else None
    L.check(lib.rgqa_gemm_bf16_reduce_logits(L.current_stream_ptr(), a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0),
                                             bias.data_ptr(), L.ptr(logits), M, N, K, flags, work.data_ptr(), work.numel(),
                                             scores.data_ptr(), labels.data_ptr()), "rgqa_gemm_bf16_reduce_logits")
    torch.cuda.synchronize()
    return scores, labels, logits


@pytest.mark.parametrize("M,N,K,scale", [(100, 1842, 1536, 0.05), (1024, 1842, 1536, 0.05), (333, 130, 64, 0.3), (2000, 1842, 1536, 0.4)])
def test_answer_head_gemm_with_scorer_epilogue(cuda_device, M, N, K, scale):
    """scores / labels out of the GEMM epilogue == the stand-alone scorer kernels run on the logits the same GEMM writes
    (same sigmoid / softplus expressions: labels and MSP bit for bit), and the logits == torch on the bf16 operands.
    scale 0.4 gives |logit| up to ~40: sigmoid saturates to 1.0f and the arg-max must still be the FIRST saturated class."""
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g)).to(torch.bfloat16).to(cuda_device)
    w = (torch.randn(N, K, generator=g) * scale / np.sqrt(K) * 8).to(torch.bfloat16).to(cuda_device)
    bias = (torch.randn(N, generator=g) * 0.5).to(cuda_device)
    scores, labels, logits = _reduce_logits(a, w, bias, 3, True)
    ref = a.float() @ w.float().t() + bias
    assert (logits - ref).abs().max() <= 2e-3 * max(1.0, float(ref.abs().max()))  # fp32 accumulation order only
    for row, mode in enumerate(("msp", "energy", "energy_sum", "lse")):
        s, l = ops.score_logits(logits, mode)
        assert torch.equal(l, labels), mode
        if mode in ("msp", "energy"):
            assert torch.equal(s, scores[row]), mode
        else:  # partial sums are merged in a different order
            torch.testing.assert_close(scores[row], s, rtol=2e-6, atol=1e-6)
    # the [M, N] matrix is optional: same scores without writing it, and without the exp / log passes of the two sums
    s2, l2, _ = _reduce_logits(a, w, bias, 0, False)
    assert torch.equal(l2, labels) and torch.equal(s2[0], scores[0]) and torch.equal(s2[1], scores[1])
    # back-to-back launches share the arrival counters: they must have been left at zero
    s3, l3, _ = _reduce_logits(a, w, bias, 3, False)
    assert torch.equal(l3, labels) and torch.equal(s3, scores)


def _args(precision, layers=(2, 1, 1)):
    return types.SimpleNamespace(llayers=layers[0], xlayers=layers[1], rlayers=layers[2], from_scratch=False, precision=precision,
                                 vocab=None, max_seq_length=20)


def _model(cuda_device, precision, layers=(2, 1, 1)):
    sys.argv = ["x"]
    from rgqa_b200.tasks.gqa_model import GQAModel_maha
    m = GQAModel_maha(1842, args=_args(precision, layers), max_seq_length=20)
    sd = syn.lxmert_state_dict(variant="rich", head_sigma=0.2, l_layers=layers[0], x_layers=layers[1], r_layers=layers[2])
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda_device)
    mu, P = syn.maha_stats()
    m.engine().set_maha(mu, P)
    return m, sd, mu, P


@pytest.mark.parametrize("B", [48, 1100])
def test_tensor_core_mahalanobis_against_fp32_path_and_literal_loop(cuda_device, B):
    """rgqa_score(MAHA) in a bf16 handle = ONE split-bf16 tcgen05 GEMM over [P; mu P] with -1/2 (q - 2G + m) and the arg-max
    in its epilogue.  Against the fp32 CUDA-core path (maha_tc = 0) on the same pooled rows and against the reference's
    literal 1842-iteration loop (oracle, gqa_maha.py:228-236): north_star's 1e-3 rel on rejection scores, identical labels."""
    from oracle import scorers_oracle as S
    from rgqa_b200 import _lib as L
    m, _, mu, P = _model(cuda_device, "bf16")
    eng = m.engine()
    g = torch.Generator().manual_seed(B)
    x = torch.tanh(torch.randn(B, 768, generator=g)).to(cuda_device)   # pooled features are tanh outputs
    s_tc, l_tc = eng.score("maha", pooled=x)
    L.load().rgqa_set_option(b"maha_tc", 0)
    try:
        s_f32, l_f32, full = eng.score("maha", pooled=x, return_full=True)
    finally:
        L.load().rgqa_set_option(b"maha_tc", 1)
    torch.cuda.synchronize()
    rel = ((s_tc - s_f32).abs() / s_f32.abs()).max().item()
    assert rel <= 1e-4, rel                       # measured ~2e-6: split-bf16 keeps 16 mantissa bits per operand
    top2 = full.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * top2[:, 0].abs()   # labels may differ only on numerical ties
    assert torch.equal(l_tc[clear], l_f32[clear]) and int(clear.sum()) >= B * 9 // 10
    n = min(B, 48)
    rs, rl, _ = S.mahalanobis(x[:n].cpu(), mu, P)
    np.testing.assert_allclose(s_tc[:n].cpu().numpy(), rs.numpy(), rtol=1e-3)
    assert (l_tc[:n].cpu() == rl).float().mean() >= 0.95 and torch.equal(l_tc[:n].cpu()[clear[:n].cpu()], rl.to(torch.int32)[clear[:n].cpu()])


@pytest.mark.parametrize("precision,B", [("fp32", 40), ("bf16", 40), ("bf16", 200)])
def test_forward_scored_equals_forward_then_score(cuda_device, precision, B):
    """rgqa_forward_scored (scorers inside the forward) against rgqa_lxmert_forward + rgqa_score on the same batch, eagerly
    and through the CUDA-graph replay (third call of a shape); B = 200 runs the deferred-LayerNorm encoder."""
    m, _, _, _ = _model(cuda_device, precision)
    eng = m.engine()
    inp = syn.make_inputs(B, 20, seed=3)
    feat, pos = inp["feat"].to(cuda_device), inp["pos"].to(cuda_device)
    ids, seg, mask = (inp[k].to(torch.int32).to(cuda_device) for k in ("ids", "seg", "mask"))
    pooled, logits = eng.forward(feat, pos, ids, seg, mask)
    modes = ["msp", "energy", "energy_sum", "lse", "maha"]
    want = {md: eng.score(md, logits=logits, pooled=pooled) for md in modes}
    for call in range(3):  # eager, capture, replay
        got, p2, l2 = eng.forward_scored(feat, pos, ids, seg, mask, modes, want_logits=True, want_pooled=True)
        torch.cuda.synchronize()
        torch.testing.assert_close(p2, pooled, rtol=0, atol=1e-6)
        torch.testing.assert_close(l2, logits, rtol=1e-5, atol=1e-5)
        for md in modes:
            s, l = got[md]
            ws, wl = want[md]
            assert (s - ws).abs().max() <= 2e-5 * max(1.0, float(ws.abs().max())), (md, call)
            assert (l == wl).float().mean() >= 0.97, (md, call)   # a different label needs a tie within the tolerance above
    # scorers only: neither matrix is written
    got, p2, l2 = eng.forward_scored(feat, pos, ids, seg, mask, ["energy", "maha"])
    assert p2 is None and l2 is None
    torch.testing.assert_close(got["energy"][0], want["energy"][0], rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(got["maha"][0], want["maha"][0], rtol=2e-5, atol=1e-5)


def test_graph_capture_counter_exposes_fresh_pointer_callers(cuda_device):
    """Large batches are captured per (shape, feature / box pointers).  A sweep over two staging slots captures twice and
    then only replays; a caller that allocates a new feature tensor per batch keeps capturing — `rgqa_graph_captures` is the
    counter that makes that cliff visible."""
    m, _, _, _ = _model(cuda_device, "bf16")
    eng = m.engine()
    B = 160
    inp = syn.make_inputs(B, 20, seed=9)
    pos = inp["pos"].to(cuda_device)
    ids, seg, mask = (inp[k].to(torch.int32).to(cuda_device) for k in ("ids", "seg", "mask"))
    slots = [inp["feat"].to(cuda_device), inp["feat"].to(cuda_device).clone()]
    base = eng.graph_captures()
    ref = None
    for i in range(8):
        pooled, _ = eng.forward(slots[i & 1], pos, ids, seg, mask, want_logits=False)
        ref = pooled if ref is None else ref
        torch.testing.assert_close(pooled, ref, rtol=0, atol=0)
    assert eng.graph_captures() - base == 2
    keep = []
    for i in range(6):  # a fresh buffer per call (kept alive so the allocator cannot hand the same pointer back)
        f = inp["feat"].to(cuda_device).clone()
        keep.append(f)
        pooled, _ = eng.forward(f, pos, ids, seg, mask, want_logits=False)
        torch.testing.assert_close(pooled, ref, rtol=0, atol=0)
    torch.cuda.synchronize()
    assert eng.graph_captures() - base == 2   # first sight of a pointer runs eagerly; nothing is captured for one-shot buffers
