"""Round-2 fused kernels on the GPU, each through its C-ABI entry point:

* `rgqa_visn_embed` — VisualFeatEncoder (src/lxrt/modeling.py:507-517) as ONE tcgen05 kernel, against torch fp32;
* `rgqa_qkv_attention_fused` — BertAttention for self-attention (:320-347) as ONE kernel, bit for bit against the two-kernel
  path (Q|K|V projection + attention) it replaces and against torch;
* the whole LXMERT / UNITER forward with both switched on vs. off (options `visn_fused`, `attn_fused`).
"""
import sys
import types

import pytest
import torch

from rgqa_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _ln(x, g, b):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps=1e-12)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,box_dim", [(100, 4), (1152, 4), (5000, 4), (333, 3)])
def test_visual_feature_encoder_in_one_kernel(cuda_device, dtype, M, box_dim):
    """(LN(feat W1^T + b1) + LN(box W2^T + b2)) / 2 as a bf16 hi/lo pair + identity statistics.  bf16 operands: exact products,
    fp32 accumulation -> 5e-5 abs on unit-variance outputs; fp32 operands go through kind::tf32 (10 mantissa bits)."""
    from rgqa_b200 import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(M + box_dim)
    K = 2048
    feat = torch.rand(M, K, generator=g)
    W1 = torch.randn(768, K, generator=g) * 0.02
    if dtype == torch.bfloat16:
        feat, W1 = feat.to(torch.bfloat16), W1.to(torch.bfloat16)
    b1 = torch.randn(768, generator=g) * 0.1
    g1, be1 = 1 + 0.1 * torch.randn(768, generator=g), 0.1 * torch.randn(768, generator=g)
    box = torch.rand(M, box_dim, generator=g) * (1.0 if box_dim == 4 else 300.0)   # also un-normalised (pixel-scale) boxes
    W2 = torch.randn(768, box_dim, generator=g) * 0.02
    b2 = torch.randn(768, generator=g) * 0.1
    g2, be2 = 1 + 0.1 * torch.randn(768, generator=g), 0.1 * torch.randn(768, generator=g)
    dev = [t.to(cuda_device).contiguous() for t in (feat, W1, b1, g1, be1, box, W2, b2, g2, be2)]
    hi = torch.empty(M, 768, dtype=torch.bfloat16, device=cuda_device)
    lo = torch.empty_like(hi)
    stats = torch.full((M, 12), float("nan"), device=cuda_device)
    L.check(lib.rgqa_visn_embed(L.current_stream_ptr(), dev[0].data_ptr(), L.DT_BF16 if dtype == torch.bfloat16 else L.DT_F32,
                                dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(), dev[4].data_ptr(), dev[5].data_ptr(), box_dim,
                                dev[6].data_ptr(), dev[7].data_ptr(), dev[8].data_ptr(), dev[9].data_ptr(), 1e-12, hi.data_ptr(),
                                lo.data_ptr(), stats.data_ptr(), M, K), "rgqa_visn_embed")
    torch.cuda.synchronize()
    x = feat.double() @ W1.double().t() + b1.double()
    y = box.double() @ W2.double().t() + b2.double()
    ref = (_ln(x, g1.double(), be1.double()) + _ln(y, g2.double(), be2.double())) / 2
    got = hi.double().cpu() + lo.double().cpu()
    err = (got - ref).abs().max().item()
    assert err <= (5e-5 if dtype == torch.bfloat16 else 4e-3), err
    assert (lo.float().abs() <= hi.float().abs() * 2.0 ** -8 + 1e-30).all()
    want = torch.zeros(12)
    want[1] = 768.0
    assert torch.equal(stats.cpu(), want.expand(M, 12))


def _hilo_state(M, g, device, scale=2.0, mean=0.2):
    x = torch.randn(M, 768, generator=g) * scale + mean
    hi = x.to(torch.bfloat16)
    v = hi.float().double()
    stats = torch.stack([v.view(M, 6, 128).sum(-1), (v * v).view(M, 6, 128).sum(-1)], dim=-1).float()
    return hi.to(device), stats.contiguous().to(device)


def _head_major(w, b):
    """[q rows | k rows | v rows] -> [head][q | k | v][64]"""
    H = 768
    idx = torch.cat([torch.arange(part * H + h * 64, part * H + h * 64 + 64) for h in range(12) for part in range(3)])
    return w[idx].contiguous(), b[idx].contiguous()


@pytest.mark.parametrize("B,S,masked,folded", [(60, 20, True, True), (1024, 20, True, True), (300, 36, False, True),
                                               (1023, 36, False, False), (33, 56, True, True), (40, 30, True, False),
                                               (20, 64, True, True), (7, 17, True, True)])
def test_qkv_projection_with_attention_epilogue(cuda_device, B, S, masked, folded):
    """One kernel == Q|K|V projection kernel + attention kernel, bit for bit (same tcgen05 accumulation, same bf16 rounding of
    the projection, same mma.sync attention arithmetic), over ragged tails (B not a multiple of the samples per tile), every
    sequence length of the path (20 / 30 language, 36 visual, 56 UNITER, 64 max) with and without the key mask, as a
    deferred-LayerNorm consumer (folded) and on already-normalised rows."""
    from rgqa_b200 import ops, _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(B * 100 + S)
    M = B * S
    hi, stats = _hilo_state(M, g, cuda_device)
    w = (torch.randn(2304, 768, generator=g) * 0.03).to(cuda_device)
    bias = (torch.randn(2304, generator=g) * 0.1).to(cuda_device)
    if folded:
        gamma = (1 + 0.1 * torch.randn(768, generator=g)).to(cuda_device)
        beta = (0.1 * torch.randn(768, generator=g)).to(cuda_device)
        wq, _, bq = ops.fold_layernorm_weight(w, bias, gamma, beta)
        qkv = ops.linear_bf16_folded(hi, stats, wq, bq, out_dtype=torch.bfloat16)
    else:
        wq, bq = w.to(torch.bfloat16), bias
        qkv = ops.linear_bf16(hi, wq, bq, out_dtype=torch.bfloat16)
    mask = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), generator=g)
        mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int32).to(cuda_device)
    want = ops.attention(qkv, qkv, B, S, S, key_mask=mask)
    w_hm, b_hm = _head_major(wq, bq)
    ctx = torch.full((M, 768), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    for _ in range(2):  # twice: the second launch reuses warm TMEM / shared memory
        L.check(lib.rgqa_qkv_attention_fused(L.current_stream_ptr(), hi.data_ptr(), 768, w_hm.data_ptr(), b_hm.data_ptr(),
                                             stats.data_ptr() if folded else None, 1e-12, L.ptr(mask), ctx.data_ptr(), 768, B, S, 12),
                "rgqa_qkv_attention_fused")
    torch.cuda.synchronize()
    assert torch.isfinite(ctx.float()).all()
    assert torch.equal(ctx, want), (ctx.float() - want.float()).abs().max().item()


def _args(precision, layers):
    return types.SimpleNamespace(llayers=layers[0], xlayers=layers[1], rlayers=layers[2], from_scratch=False, precision=precision,
                                 vocab=None, max_seq_length=20)


@pytest.mark.parametrize("B", [64, 250])
def test_lxmert_forward_fused_kernels_against_separate_kernels(cuda_device, B):
    """Whole bf16 forward (deferred LayerNorm, B * 20 >= 1024 rows) with the one-kernel visual encoder and the attention
    epilogue on vs. off: the attention fusion is bit-identical, the visual encoder differs by its tf32 (instead of bf16)
    read of the features — inside the bf16 path's own noise (rms 5e-3 on logits of std 0.79)."""
    sys.argv = ["x"]
    from rgqa_b200 import _lib as L
    from rgqa_b200.tasks.gqa_model import GQAModel_maha
    lib = L.load()
    layers = (2, 2, 2)
    m = GQAModel_maha(1842, args=_args("bf16", layers), max_seq_length=20)
    m.load_state_dict(syn.lxmert_state_dict(variant="rich", l_layers=2, x_layers=2, r_layers=2), strict=True)
    m = m.to(cuda_device)
    eng = m.engine()
    inp = syn.make_inputs(B, 20, seed=21)
    feat, pos = inp["feat"].to(cuda_device), inp["pos"].to(cuda_device)
    ids, seg, mask = (inp[k].to(torch.int32).to(cuda_device) for k in ("ids", "seg", "mask"))
    out = {}
    try:
        for name, (vf, af) in {"off": (0, 0), "attn": (0, 1), "visn": (1, 0), "both": (1, 1)}.items():
            lib.rgqa_set_option(b"visn_fused", vf)
            lib.rgqa_set_option(b"attn_fused", af)
            out[name] = eng.forward(feat, pos, ids, seg, mask)
            torch.cuda.synchronize()
    finally:
        lib.rgqa_set_option(b"visn_fused", 1)
        lib.rgqa_set_option(b"attn_fused", 1)
    assert torch.equal(out["attn"][1], out["off"][1]) and torch.equal(out["attn"][0], out["off"][0])
    assert torch.equal(out["both"][1], out["visn"][1])
    d = (out["visn"][1] - out["off"][1])   # two bf16-noise paths against each other (measured: max 1.6e-2, rms 3.3e-3)
    assert d.abs().max().item() <= 2.5e-2 and d.pow(2).mean().sqrt().item() <= 5e-3, (d.abs().max().item(), d.pow(2).mean().sqrt().item())
    # bf16 features (the ingest store's format) take the kind::f16 variant of the same kernel
    pooled_b, logits_b = eng.forward(feat.to(torch.bfloat16), pos, ids, seg, mask)
    d = (logits_b - out["off"][1])
    assert d.abs().max().item() <= 2.5e-2 and d.pow(2).mean().sqrt().item() <= 5e-3


def test_uniter_forward_attention_epilogue_bit_identical(cuda_device):
    """UNITER's 12 single-stream layers (S = 56 joint tokens, src/uniter/modeling.py:623-635) through the fused projection +
    attention kernel == the separate kernels."""
    from rgqa_b200 import _lib as L
    from rgqa_b200.engine import Engine, EngineConfig
    lib = L.load()
    cfg = EngineConfig(arch="uniter", precision="bf16", l_layers=3, x_layers=0, r_layers=0, vocab_size=28996, visual_pos_dim=7)
    eng = Engine(cfg)
    eng.load_weights(syn.uniter_state_dict(variant="rich", layers=3).items())
    B = 40
    inp = syn.make_inputs(B, 20, seed=5, vocab=28996)
    feat, pos = inp["feat"].to(cuda_device), inp["pos7"].to(cuda_device)
    ids, seg, mask = (inp[k].to(torch.int32).to(cuda_device) for k in ("ids", "seg", "mask"))
    try:
        lib.rgqa_set_option(b"attn_fused", 0)
        p0, l0 = eng.forward(feat, pos, ids, seg, mask)
        lib.rgqa_set_option(b"attn_fused", 1)
        p1, l1 = eng.forward(feat, pos, ids, seg, mask)
        torch.cuda.synchronize()
    finally:
        lib.rgqa_set_option(b"attn_fused", 1)
    assert torch.equal(p0, p1) and torch.equal(l0, l1)


@pytest.mark.parametrize("attn_fused", [0, 1])
def test_programmatic_dependent_launch_is_bit_identical(cuda_device, attn_fused):
    """Option `pdl` launches the hot kernels with the programmatic-stream-serialization attribute: each kernel's prologue may
    then overlap its predecessor's tail, and every such kernel must `griddepcontrol.wait` before its first global access.
    ADVICE r01 found `attention_tma_kernel` launched that way without the wait (its producer warp could TMA-load a projection
    still being written).  With the wait in place the whole forward — TMA attention (attn_fused = 0: all 34 blocks; = 1: the 10
    cross blocks) — is bit-identical with and without PDL."""
    sys.argv = ["x"]
    from rgqa_b200 import _lib as L
    from rgqa_b200.engine import Engine, EngineConfig
    lib = L.load()
    cfg = EngineConfig(precision="bf16", l_layers=2, x_layers=2, r_layers=2)
    eng = Engine(cfg)
    eng.load_weights(syn.lxmert_state_dict(variant="rich", l_layers=2, x_layers=2, r_layers=2).items())
    B = 260   # 260 x 12 units: above the TMA attention kernel's threshold
    inp = syn.make_inputs(B, 20, seed=77)
    t = [inp["feat"].to(cuda_device), inp["pos"].to(cuda_device)] + [inp[k].to(torch.int32).to(cuda_device) for k in ("ids", "seg", "mask")]
    out = []
    try:
        lib.rgqa_set_option(b"graphs", 0)
        lib.rgqa_set_option(b"attn_fused", attn_fused)
        for pdl in (0, 1, 1):
            lib.rgqa_set_option(b"pdl", pdl)
            out.append(eng.forward(*t))
            torch.cuda.synchronize()
    finally:
        lib.rgqa_set_option(b"pdl", 0)
        lib.rgqa_set_option(b"graphs", 1)
        lib.rgqa_set_option(b"attn_fused", 1)
    eng.close()
    for pooled, logits in out[1:]:
        assert torch.equal(pooled, out[0][0]) and torch.equal(logits, out[0][1])


@pytest.mark.parametrize("B,masked,f16", [(29, False, False), (35, True, True), (300, False, True), (1023, True, False), (1024, False, True),
                                          (1030, True, True)])
def test_seven_samples_per_cta_pair_equals_three_per_cta(cuda_device, B, masked, f16):
    """S = 36 (visual stream): the straddle form (seven whole samples per CTA pair, the middle sample's K / V lines crossing the pair
    through distributed shared memory, its query tiles split 1 + 2 between the CTAs) is bit-identical to the 3-samples-per-CTA form
    and to the separate projection + attention kernels — every residue of B mod 7 (tail tiles with 1..6 samples, with and without
    a straddling sample), bf16 and fp16 operands, several launches back to back (the halo buffers' hand-shake across tiles)."""
    from rgqa_b200 import ops, _lib as L
    lib = L.load()
    S = 36
    g = torch.Generator().manual_seed(B)
    M = B * S
    x = torch.randn(M, 768, generator=g) * 2.0 + 0.2
    hi = x.to(torch.float16 if f16 else torch.bfloat16)
    v = hi.double()
    stats = torch.stack([v.view(M, 6, 128).sum(-1), (v * v).view(M, 6, 128).sum(-1)], dim=-1).float().contiguous().to(cuda_device)
    hi = hi.to(cuda_device)
    w = (torch.randn(2304, 768, generator=g) * 0.03).to(cuda_device)
    bias = (torch.randn(2304, generator=g) * 0.1).to(cuda_device)
    gamma = (1 + 0.1 * torch.randn(768, generator=g)).to(cuda_device)
    beta = (0.1 * torch.randn(768, generator=g)).to(cuda_device)
    wq, _, bq = ops.fold_layernorm_weight(w, bias, gamma, beta, dtype=torch.float16 if f16 else torch.bfloat16)
    qkv = (ops.linear_a16_folded if f16 else ops.linear_bf16_folded)(hi, stats, wq, bq, out_dtype=torch.bfloat16)
    mask = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), generator=g)
        mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int32).to(cuda_device)
    want = ops.attention(qkv, qkv, B, S, S, key_mask=mask)
    w_hm, b_hm = _head_major(wq, bq)
    fn = lib.rgqa_qkv_attention_fused_a16 if f16 else lib.rgqa_qkv_attention_fused
    got = {}
    for straddle in (1, 0):
        L.check(lib.rgqa_set_option(b"attn_straddle", straddle), "opt")
        try:
            ctx = torch.full((M + 64, 768), float("nan"), dtype=torch.bfloat16, device=cuda_device)   # guard rows behind the last sample
            for _ in range(3):
                L.check(fn(L.current_stream_ptr(), hi.data_ptr(), 768, w_hm.data_ptr(), b_hm.data_ptr(), stats.data_ptr(), 1e-12, L.ptr(mask),
                           ctx.data_ptr(), 768, B, S, 12), "fused")
            torch.cuda.synchronize()
        finally:
            L.check(lib.rgqa_set_option(b"attn_straddle", 0), "opt")
        assert torch.isnan(ctx[M:].float()).all()           # nothing written behind the last sample's rows
        got[straddle] = ctx[:M]
        assert torch.isfinite(got[straddle].float()).all()
    assert torch.equal(got[1], got[0]), (got[1].float() - got[0].float()).abs().max().item()
    assert torch.equal(got[1], want), (got[1].float() - want.float()).abs().max().item()
