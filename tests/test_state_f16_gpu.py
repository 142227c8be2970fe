"""The single-word (fp16) deferred-LayerNorm state — option "state_f16", the default of the bf16 forward:

* tcgen05 kind::f16 with fp16 operands (the consumers of the state run fp16 x fp16 against fp16 copies of their weights — the
  hardware rejects fp16 x bf16), pair kernel and single-CTA kernel;
* `rgqa_gemm_bf16_resid_h16` — BertAttOutput / BertOutput (src/lxrt/modeling.py:357-361, :411-415) up to the LayerNorm, written
  as ONE fp16 word per element + the statistics of the rounded row;
* the whole forward with the fp16 state against the bf16 hi/lo pair and against the fp32 CUDA path."""
import sys
import types

import pytest
import torch

from rgqa_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _gelu_erf(x):
    return x * 0.5 * (1.0 + torch.erf(x / 2.0 ** 0.5))


def _h16_state(M, g, device, mean=0.2, scale=1.5):
    """(x f64 = the fp16 values, h fp16, stats [M,6,2]) for random pre-LayerNorm rows."""
    h = (torch.randn(M, 768, generator=g) * scale + mean).to(torch.float16)
    v = h.double()
    stats = torch.stack([v.view(M, 6, 128).sum(-1), (v * v).view(M, 6, 128).sum(-1)], dim=-1).float()
    return v, h.to(device), stats.contiguous().to(device)


@pytest.mark.parametrize("M,N,epi", [(2048, 2304, "bias"), (1500, 3072, "gelu"), (300, 768, "bias"), (96, 768, "tanh"), (4096, 768, "bias")])
def test_fp16_operands_on_the_tensor_cores(cuda_device, M, N, epi):
    """out = epilogue(A·Wᵀ + b) with A and W fp16: the tensor cores must multiply the operands AS GIVEN (fp32 accumulate), so
    the result equals the fp64 product of the exact operand values up to the summation order — reading the fp16 bits as bf16
    would be off by orders of magnitude, a hidden conversion to bf16 by ~5e-3."""
    from rgqa_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, 768, generator=g) * 1.3 + 0.1).to(torch.float16).to(cuda_device)
    w = (torch.randn(N, 768, generator=g) * 0.03).to(torch.float16).to(cuda_device)
    bias = (torch.randn(N, generator=g) * 0.1).to(cuda_device)
    code = {"bias": L.EPI_BIAS, "gelu": L.EPI_GELU, "tanh": L.EPI_TANH}[epi]
    out = ops.linear_a16_folded(a, None, w, bias, epilogue=code, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t() + bias.double()
    ref = _gelu_erf(ref) if epi == "gelu" else (torch.tanh(ref) if epi == "tanh" else ref)
    err = (out.double() - ref).abs().max().item()
    assert err <= 2e-4, err      # fp32 accumulation over K = 768 of O(1) sums
    outb = ops.linear_a16_folded(a, None, w, bias, epilogue=code, out_dtype=torch.bfloat16)
    assert torch.equal(outb, out.to(torch.bfloat16))


@pytest.mark.parametrize("M,N,epi", [(2048, 2304, "bias"), (1500, 3072, "gelu"), (300, 768, "bias")])
def test_folded_layernorm_consumer_of_the_fp16_state(cuda_device, M, N, epi):
    """LN(x)·Wᵀ + b from the un-normalised fp16 rows + row statistics, gamma folded into the fp16 copy of the weight."""
    from rgqa_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(M + N + 1)
    v, h, stats = _h16_state(M, g, cuda_device)
    w = (torch.randn(N, 768, generator=g) * 0.03).to(cuda_device)
    bias = (torch.randn(N, generator=g) * 0.1).to(cuda_device)
    gamma = (1 + 0.1 * torch.randn(768, generator=g)).to(cuda_device)
    beta = (0.1 * torch.randn(768, generator=g)).to(cuda_device)
    wf, c, b2 = ops.fold_layernorm_weight(w, bias, gamma, beta, dtype=torch.float16)
    wg = w.double() * gamma.double()
    assert (wf.double() - (wg - wg.mean(1, keepdim=True))).abs().max().item() <= 2.0 ** -11 * 1.01 * wg.abs().max().item()   # one fp16 ulp
    assert c.abs().max().item() <= 1e-6     # error-feedback rounding: the row sums carry one element's rounding error
    code = {"bias": L.EPI_BIAS, "gelu": L.EPI_GELU}[epi]
    out = ops.linear_a16_folded(h, stats, wf, b2, epilogue=code, out_dtype=torch.float32)
    torch.cuda.synchronize()
    y = torch.nn.functional.layer_norm(v.to(cuda_device), (768,), gamma.double(), beta.double(), eps=1e-12)
    ref = y @ w.double().t() + bias.double()
    if epi == "gelu":
        ref = _gelu_erf(ref)
    err = (out.double() - ref).abs()
    # 11-bit operands on both sides: an eighth of the bf16 consumer's error (3e-2 / 5e-3 in test_kernels_gpu.py)
    assert err.max().item() <= 5e-3 and err.pow(2).mean().sqrt().item() <= 8e-4, (err.max().item(), err.pow(2).mean().sqrt().item())
    # exact statement of what the kernel computes: rstd * (x · W'ᵀ) + b' over the operand values as given
    mu = v.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt((v * v).mean(1, keepdim=True) - mu * mu + 1e-12)
    exact = rstd.to(cuda_device) * (v.to(cuda_device) @ wf.double().t()) + b2.double()
    exact = _gelu_erf(exact) if epi == "gelu" else exact
    assert (out.double() - exact).abs().max().item() <= 5e-4


@pytest.mark.parametrize("fold_beta", [False, True])
@pytest.mark.parametrize("M,K", [(1024, 768), (2048, 3072), (1300, 768), (36864, 768)])
def test_gemm_residual_fp16_state_producer(cuda_device, M, K, fold_beta):
    """v = A·Wᵀ + b + LN_prev(x_prev) written as ONE fp16 word per element + the partial statistics of the rounded row."""
    from rgqa_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    v_in, h_in, st_in = _h16_state(M, g, cuda_device)
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    w = (torch.randn(768, K, generator=g) * 0.03).to(torch.bfloat16).to(cuda_device)
    bias = (torch.randn(768, generator=g) * 0.1).to(cuda_device)
    gamma = (1 + 0.1 * torch.randn(768, generator=g)).to(cuda_device)
    beta = (0.1 * torch.randn(768, generator=g)).to(cuda_device)
    if fold_beta:
        h, stats = ops.linear_residual_h16(a, w, bias + beta, h_in, st_in, gamma, None)
    else:
        h, stats = ops.linear_residual_h16(a, w, bias, h_in, st_in, gamma, beta)
    torch.cuda.synchronize()
    resid = torch.nn.functional.layer_norm(v_in.to(cuda_device), (768,), gamma.double(), beta.double(), eps=1e-12)
    ref = a.double() @ w.double().t() + bias.double() + resid
    got = h.double()
    # one fp16 rounding of |v| <~ 8: half an ulp = 2^-11 relative, plus the fp32 accumulation noise
    assert ((got - ref).abs() <= ref.abs() * 2.0 ** -11 + 1e-3).all(), (got - ref).abs().max().item()
    want = torch.stack([got.view(M, 6, 128).sum(-1), (got * got).view(M, 6, 128).sum(-1)], dim=-1)
    torch.testing.assert_close(stats.double(), want, atol=2e-2, rtol=2e-4)


def test_fp16_state_saturates_instead_of_overflowing(cuda_device):
    """A pre-LayerNorm value beyond the fp16 range is stored as +-65504 (never inf / nan): the next LayerNorm then sees a large
    finite outlier, as it would in fp32."""
    from rgqa_b200 import ops
    M, K = 1024, 768
    g = torch.Generator().manual_seed(5)
    v_in, h_in, st_in = _h16_state(M, g, cuda_device)
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    w = (torch.randn(768, K, generator=g) * 0.03).to(torch.bfloat16).to(cuda_device)
    bias = torch.zeros(768)
    bias[5], bias[700] = 1e6, -3e5
    gamma = torch.ones(768, device=cuda_device)
    h, stats = ops.linear_residual_h16(a, w, bias.to(cuda_device), h_in, st_in, gamma, None)
    torch.cuda.synchronize()
    assert torch.isfinite(h.float()).all() and torch.isfinite(stats).all()
    assert (h[:, 5].float() == 65504.0).all() and (h[:, 700].float() == -65504.0).all()


def _args(precision, layers):
    return types.SimpleNamespace(llayers=layers[0], xlayers=layers[1], rlayers=layers[2], from_scratch=False, precision=precision,
                                 vocab=None, max_seq_length=20)


@pytest.mark.parametrize("variant", ["init", "rich"])
def test_forward_with_the_fp16_state_against_the_hilo_pair_and_fp32(cuda_device, variant):
    """Full-depth 9/5/5 forward, B = 256: the fp16 single-word state must be at least as close to the fp32 CUDA path as the
    bf16 hi/lo pair it replaces (the consumers' A operand carries 11 instead of 8 significand bits; the residual carrier 11
    instead of 16 — the second is below the first's noise), and inside the bf16 path's stated bounds."""
    sys.argv = ["x"]
    from rgqa_b200 import _lib as L
    from rgqa_b200.tasks.gqa_model import GQAModel_maha
    lib = L.load()
    sd = syn.lxmert_state_dict(seed=77, variant=variant)
    inp = syn.make_inputs(256, 20, seed=4)
    outs = {}
    args = [inp[k].to(cuda_device) for k in ("feat", "pos", "ids", "seg", "mask")]
    for name, prec, st in (("fp32", "fp32", 1), ("bf16_state1", "bf16", 1), ("bf16_state0", "bf16", 0)):
        L.check(lib.rgqa_set_option(b"state_f16", st), "opt")     # read when the weights are finalized (model.to(device))
        try:
            m = GQAModel_maha(1842, args=_args(prec, (9, 5, 5)), max_seq_length=20)
            m.load_state_dict(sd, strict=True)
            m = m.to(cuda_device)
            outs[name] = m.forward_ids(*args)[0].float().clone()
            torch.cuda.synchronize()
        finally:
            L.check(lib.rgqa_set_option(b"state_f16", 1), "opt")
    ref = outs["fp32"]
    scale = ref.std().item() / 0.79
    e16 = (outs["bf16_state1"] - ref)
    ehl = (outs["bf16_state0"] - ref)
    rms16, rmshl = e16.pow(2).mean().sqrt().item(), ehl.pow(2).mean().sqrt().item()
    print(f"STATE_F16 variant={variant} rms fp16-state {rms16:.5f} max {e16.abs().max().item():.4f} | hi/lo pair rms {rmshl:.5f} max {ehl.abs().max().item():.4f}")
    assert rms16 <= 6e-3 * scale and e16.abs().max().item() <= 3e-2 * scale, (rms16, e16.abs().max().item(), scale)
    assert rms16 <= 1.05 * rmshl, (rms16, rmshl)
