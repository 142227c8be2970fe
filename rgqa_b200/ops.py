"""Tensor-level wrappers over the per-kernel C-ABI entry points.

Each function borrows the data pointers of CUDA ``torch.Tensor`` arguments, enqueues ONE hand-written
kernel on the current CUDA stream and returns a fresh output tensor.  They mirror single modules of
the reference (file:line in each docstring) so that unit tests read like ``module(x)`` comparisons.
PyTorch is only the allocator / stream provider here.
"""
from __future__ import annotations

import torch

from . import _lib as L


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("rgqa_b200.ops: CUDA tensors required (there is no CPU path)")


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.DT_F32
    if t.dtype == torch.bfloat16:
        return L.DT_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def linear_bf16(x, weight, bias=None, epilogue=L.EPI_BIAS, residual=None, out_dtype=torch.bfloat16, out=None):
    """``epilogue(x @ weight.T + bias [+ residual])`` on the tcgen05 tensor cores.

    x: [M, K] bf16 (row stride may exceed K), weight: [N, K] bf16 (nn.Linear layout), bias: [N] f32,
    residual: [M, N] f32.
    Reference: every ``nn.Linear`` of src/lxrt/modeling.py (e.g. :321-323, :358, :399, :412, :578).
    """
    _need_cuda(x, weight, bias, residual)
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.dim() == 2 and weight.dim() == 2
    assert x.stride(1) == 1 and weight.stride(1) == 1
    M, K = x.shape
    N = weight.shape[0]
    assert weight.shape[1] == K
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    else:  # caller-provided destination (may alias `residual`: an accumulating update)
        assert out.shape == (M, N) and out.stride(1) == 1 and out.is_cuda
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    ldr = 0
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1  # the fp32 residual stream
        ldr = residual.stride(0)
    lib = L.load()
    L.check(lib.rgqa_gemm_bf16(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(weight), weight.stride(0),
                               L.ptr(bias), L.ptr(residual), ldr, L.ptr(out), out.stride(0), _dt(out), M, N, K,
                               epilogue), "rgqa_gemm_bf16")
    return out


def linear_residual_layernorm_bf16(x, weight, bias, residual, gamma, beta, eps=1e-12, want_f32=True):
    """(bf16 [M,768], f32 [M,768]) = LayerNorm(x @ weight.T + bias + residual) — BertAttOutput / BertOutput
    (src/lxrt/modeling.py:357-361, :411-415) in one tcgen05 kernel; x bf16 [M,K], weight bf16 [768,K], residual f32."""
    _need_cuda(x, weight, bias, residual, gamma, beta)
    M, K = x.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    out32 = torch.empty((M, N), dtype=torch.float32, device=x.device) if want_f32 else None
    lib = L.load()
    L.check(lib.rgqa_gemm_bf16_ln(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(weight), weight.stride(0), L.ptr(bias),
                                  L.ptr(residual), residual.stride(0), L.ptr(gamma), L.ptr(beta), eps, L.ptr(out), L.ptr(out32),
                                  M, N, K), "rgqa_gemm_bf16_ln")
    return out, out32


def fold_layernorm_weight(weight, bias, gamma, beta, dtype=torch.bfloat16):
    """(W' bf16 [N,K], c f32 [N], bias' f32 [N]) with the LayerNorm affine of the *input* folded into an nn.Linear:
    LN(x)·Wᵀ + b = rstd·(x·W'ᵀ − mean·c) + bias', W' = bf16(W∘gamma − rowmean(W∘gamma)), c = rowsum(W') ≈ 0.
    ``dtype=torch.float16``: W' rounded to fp16 instead (consumers of the single-word fp16 state)."""
    _need_cuda(weight, bias, gamma, beta)
    N, K = weight.shape
    if dtype == torch.float16:
        wf = torch.empty((N, K), dtype=torch.float16, device=weight.device)
        c = torch.empty(N, dtype=torch.float32, device=weight.device)
        b2 = torch.empty(N, dtype=torch.float32, device=weight.device)
        L.check(L.load().rgqa_fold_ln_weight_f16(L.current_stream_ptr(), L.ptr(weight), L.ptr(bias), L.ptr(gamma), L.ptr(beta), N, K,
                                                 L.ptr(wf), L.ptr(c), L.ptr(b2)), "rgqa_fold_ln_weight_f16")
        return wf, c, b2
    wf = torch.empty((N, K), dtype=torch.bfloat16, device=weight.device)
    c = torch.empty(N, dtype=torch.float32, device=weight.device)
    b2 = torch.empty(N, dtype=torch.float32, device=weight.device)
    lib = L.load()
    L.check(lib.rgqa_fold_ln_weight(L.current_stream_ptr(), L.ptr(weight), L.ptr(bias), L.ptr(gamma), L.ptr(beta), N, K,
                                    L.ptr(wf), L.ptr(c), L.ptr(b2)), "rgqa_fold_ln_weight")
    return wf, c, b2


def linear_bf16_folded(x_hi, stats, w_folded, bias_folded, c=None, epilogue=L.EPI_BIAS, residual=None, out_dtype=torch.bfloat16,
                       eps=1e-12):
    """Consumer of a deferred-LayerNorm hidden state: epilogue(LN(x)·Wᵀ + b) from the un-normalised bf16 rows ``x_hi``,
    their partial statistics ``stats`` [M,6,2] and the folded weight of :func:`fold_layernorm_weight`."""
    _need_cuda(x_hi, stats, w_folded, c, bias_folded, residual)
    M, K = x_hi.shape
    N = w_folded.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=x_hi.device)
    ldr = residual.stride(0) if residual is not None else 0
    lib = L.load()
    L.check(lib.rgqa_gemm_bf16_folded(L.current_stream_ptr(), L.ptr(x_hi), x_hi.stride(0), L.ptr(w_folded), w_folded.stride(0),
                                      L.ptr(bias_folded), L.ptr(residual), ldr, L.ptr(out), out.stride(0), _dt(out), M, N, K,
                                      epilogue, L.ptr(stats), L.ptr(c), eps), "rgqa_gemm_bf16_folded")
    return out


def linear_residual_hilo(x, weight, bias, in_hi, in_lo, in_stats, in_gamma, in_beta, eps=1e-12):
    """Producer of a deferred-LayerNorm hidden state: v = x·Wᵀ + b + LN_prev(in_hi + in_lo); returns (hi, lo, stats)
    with hi = bf16(v), lo = bf16(v − hi), stats [M,6,2] = per-128-column (sum, sumsq) of hi + lo."""
    _need_cuda(x, weight, bias, in_hi, in_lo, in_stats, in_gamma, in_beta)
    M, K = x.shape
    N = weight.shape[0]
    hi = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    stats = torch.empty((M, 6, 2), dtype=torch.float32, device=x.device)
    lib = L.load()
    L.check(lib.rgqa_gemm_bf16_resid_hl(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(weight), weight.stride(0),
                                        L.ptr(bias), L.ptr(in_hi), L.ptr(in_lo), L.ptr(in_stats), L.ptr(in_gamma),
                                        L.ptr(in_beta), eps, L.ptr(hi), L.ptr(lo), L.ptr(stats), M, N, K),
            "rgqa_gemm_bf16_resid_hl")
    return hi, lo, stats


def linear_a16_folded(x_h, stats, w_folded, bias_folded, c=None, epilogue=L.EPI_BIAS, out_dtype=torch.bfloat16, eps=1e-12):
    """:func:`linear_bf16_folded` over the single-word state: ``x_h`` fp16 [M,768] (A operand) against the fp16 folded weight;
    ``stats`` None = plain epilogue(x_h·Wᵀ + b) (the identity state in front of the first layer)."""
    _need_cuda(x_h, stats, w_folded, c, bias_folded)
    assert x_h.dtype == torch.float16 and w_folded.dtype == torch.float16
    M, K = x_h.shape
    N = w_folded.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=x_h.device)
    lib = L.load()
    L.check(lib.rgqa_gemm_bf16_folded_a16(L.current_stream_ptr(), L.ptr(x_h), x_h.stride(0), L.ptr(w_folded), w_folded.stride(0),
                                          L.ptr(bias_folded), L.ptr(out), out.stride(0), _dt(out), M, N, K, epilogue,
                                          L.ptr(stats), L.ptr(c), eps), "rgqa_gemm_bf16_folded_a16")
    return out


def linear_residual_h16(x, weight, bias, in_h, in_stats, in_gamma, in_beta, eps=1e-12):
    """Producer of the single-word deferred-LayerNorm state: v = x·Wᵀ + b + LN_prev(in_h); returns (h, stats) with
    h = fp16(v) (saturating) and stats [M,6,2] = per-128-column (sum, sumsq) of the rounded row."""
    _need_cuda(x, weight, bias, in_h, in_stats, in_gamma, in_beta)
    assert in_h.dtype == torch.float16
    M, K = x.shape
    N = weight.shape[0]
    h = torch.empty((M, N), dtype=torch.float16, device=x.device)
    stats = torch.empty((M, 6, 2), dtype=torch.float32, device=x.device)
    lib = L.load()
    L.check(lib.rgqa_gemm_bf16_resid_h16(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(weight), weight.stride(0),
                                         L.ptr(bias), L.ptr(in_h), L.ptr(in_stats), L.ptr(in_gamma), L.ptr(in_beta), eps,
                                         L.ptr(h), L.ptr(stats), M, N, K), "rgqa_gemm_bf16_resid_h16")
    return h, stats


def linear_bf16_accumulate_(out, x, weight):
    """out [M, N] f32 += x [M, K] · weight [N, K]ᵀ (bf16 operands, tensor cores, K split over CTAs, atomic accumulation): the
    weight-gradient product of the mixed-precision training step.  In place."""
    _need_cuda(out, x, weight)
    assert x.dtype == weight.dtype == torch.bfloat16 and out.dtype == torch.float32
    assert x.stride(1) == 1 and weight.stride(1) == 1 and out.stride(1) == 1
    M, K = x.shape
    N = weight.shape[0]
    assert out.shape == (M, N) and weight.shape[1] == K
    L.check(L.load().rgqa_gemm_bf16_accum(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(weight), weight.stride(0), L.ptr(out),
                                          out.stride(0), M, N, K), "rgqa_gemm_bf16_accum")
    return out


def linear_bf16_accumulate_tn_(out, at, bt):
    """out [M, N] f32 += atᵀ · bt for bf16 at [K, M], bt [K, N] (row strides allowed): the weight gradient dW += dYᵀ · X straight
    from the rows of dY and X — the tensor cores read both operands MN-major, nothing is transposed in memory.  In place."""
    _need_cuda(out, at, bt)
    assert at.dtype == bt.dtype == torch.bfloat16 and out.dtype == torch.float32
    assert at.stride(1) == 1 and bt.stride(1) == 1 and out.stride(1) == 1 and at.shape[0] == bt.shape[0]
    K, M = at.shape
    N = bt.shape[1]
    assert out.shape == (M, N)
    L.check(L.load().rgqa_gemm_bf16_accum_tn(L.current_stream_ptr(), L.ptr(at), at.stride(0), L.ptr(bt), bt.stride(0), L.ptr(out),
                                             out.stride(0), M, N, K), "rgqa_gemm_bf16_accum_tn")
    return out


def cast_bf16_colsum(x, colsum_into=None, out=None):
    """bf16 copy of an fp32 [R, C] view (row stride allowed) — dense, or into ``out`` (a bf16 [R, C] view, e.g. a column block of a
    wider buffer); ``colsum_into`` [C] f32 += the column sums of x (same pass)."""
    _need_cuda(x, colsum_into, out)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    if out is None:
        out = torch.empty((R, C), dtype=torch.bfloat16, device=x.device)
    assert out.dtype == torch.bfloat16 and out.shape == (R, C) and out.stride(1) == 1
    L.check(L.load().rgqa_cast_bf16_colsum(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(out), out.stride(0), R, C,
                                           L.ptr(colsum_into)), "rgqa_cast_bf16_colsum")
    return out


def linear_f32(x, weight, bias=None, epilogue=L.EPI_BIAS, residual=None):
    """fp32 CUDA-core GEMM with the same contract as :func:`linear_bf16` (precise path)."""
    _need_cuda(x, weight, bias, residual)
    assert x.dtype == torch.float32 and weight.dtype == torch.float32
    assert x.stride(1) == 1 and weight.stride(1) == 1
    M, K = x.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    ldr = residual.stride(0) if residual is not None else 0
    lib = L.load()
    L.check(lib.rgqa_gemm_f32(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(weight), weight.stride(0),
                              L.ptr(bias), L.ptr(residual), ldr, L.ptr(out), out.stride(0), M, N, K, epilogue),
            "rgqa_gemm_f32")
    return out


def attention(qkv_q, qkv_kv, B, Lq, Lk, key_mask=None, heads=12):
    """Multi-head attention reading Q from ``qkv_q`` [B*Lq, 3H] and K, V from ``qkv_kv`` [B*Lk, 3H]
    (fused Q|K|V projections, H = heads*64).  key_mask: int32 [B, Lk] (1 = keep) or None.
    Reference: BertAttention.forward, src/lxrt/modeling.py:320-347."""
    _need_cuda(qkv_q, qkv_kv, key_mask)
    H = heads * 64
    assert qkv_q.shape == (B * Lq, 3 * H) and qkv_kv.shape == (B * Lk, 3 * H)
    assert qkv_q.dtype == qkv_kv.dtype and qkv_q.is_contiguous() and qkv_kv.is_contiguous()
    out = torch.empty((B * Lq, H), dtype=qkv_q.dtype, device=qkv_q.device)
    es = qkv_q.element_size()
    if key_mask is not None:
        assert key_mask.dtype == torch.int32 and key_mask.shape == (B, Lk) and key_mask.is_contiguous()
    lib = L.load()
    L.check(lib.rgqa_attention(L.current_stream_ptr(), _dt(qkv_q), L.ptr(qkv_q), 3 * H,
                               qkv_kv.data_ptr() + H * es, 3 * H, qkv_kv.data_ptr() + 2 * H * es, 3 * H,
                               L.ptr(key_mask), L.ptr(out), H, B, Lq, Lk, heads), "rgqa_attention")
    return out


def embed_ln(ids, seg, word, pos, type_, gamma, beta, eps=1e-12, out_dtype=torch.float32):
    """BertEmbeddings.forward (src/lxrt/modeling.py:278-292): ids/seg int32 [B, L] -> [B*L, H]."""
    _need_cuda(ids, seg, word, pos, type_, gamma, beta)
    assert ids.dtype == torch.int32 and ids.is_contiguous()
    B, Lq = ids.shape
    H = word.shape[1]
    out = torch.empty((B * Lq, H), dtype=out_dtype, device=ids.device)
    lib = L.load()
    L.check(lib.rgqa_embed_ln(L.current_stream_ptr(), _dt(out), L.ptr(ids), L.ptr(seg), L.ptr(word), L.ptr(pos),
                              L.ptr(type_), L.ptr(gamma), L.ptr(beta), L.ptr(out), B, Lq, H, word.shape[0],
                              pos.shape[0], type_.shape[0], eps), "rgqa_embed_ln")
    return out


def layernorm(x, gamma, beta, eps=1e-12, out_dtype=torch.float32):
    """torch.nn.LayerNorm over the last dim of fp32 [M, H] (H in {768, 1536})."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    M, H = x.shape
    out = torch.empty((M, H), dtype=out_dtype, device=x.device)
    lib = L.load()
    L.check(lib.rgqa_layernorm(L.current_stream_ptr(), _dt(out), L.ptr(x), x.stride(0), L.ptr(gamma), L.ptr(beta),
                               L.ptr(out), M, H, eps), "rgqa_layernorm")
    return out


def layernorm_dual(x, gamma, beta, eps=1e-12):
    """(fp32 y, bf16 y) = LayerNorm(x) written twice by one kernel: the fp32 rows are the residual stream / the tape, the bf16
    rows the next GEMM's operand (mixed-precision training step)."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    M, H = x.shape
    y32 = torch.empty((M, H), dtype=torch.float32, device=x.device)
    y16 = torch.empty((M, H), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().rgqa_layernorm_dual(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(gamma), L.ptr(beta), L.ptr(y16), L.ptr(y32),
                                         M, H, eps), "rgqa_layernorm_dual")
    return y32, y16


def visn_combine(x_pre, boxes, box_w, box_b, g1, b1, g2, b2, eps=1e-12, out_dtype=torch.float32):
    """Tail of VisualFeatEncoder.forward (src/lxrt/modeling.py:507-517): (LN(x_pre) + LN(box_fc(boxes))) / 2."""
    _need_cuda(x_pre, boxes, box_w, box_b, g1, b1, g2, b2)
    M, H = x_pre.shape
    boxes = boxes.reshape(M, -1).contiguous()
    out = torch.empty((M, H), dtype=out_dtype, device=x_pre.device)
    lib = L.load()
    L.check(lib.rgqa_visn_combine(L.current_stream_ptr(), _dt(out), L.ptr(x_pre), L.ptr(boxes), boxes.shape[1],
                                  L.ptr(box_w), L.ptr(box_b), L.ptr(g1), L.ptr(b1), L.ptr(g2), L.ptr(b2), L.ptr(out),
                                  M, H, eps), "rgqa_visn_combine")
    return out


def cast_bf16(x):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lib = L.load()
    L.check(lib.rgqa_cast_f32_to_bf16(L.current_stream_ptr(), L.ptr(x), L.ptr(out), x.numel()), "rgqa_cast_f32_to_bf16")
    return out


def transpose_cast_bf16(x, colsum_into=None):
    """bf16 [C, R] = x.t() for fp32 x [R, C] (any row stride): transposing copy + cast in one kernel.  ``colsum_into`` [C] f32:
    the column sums of x are ADDED to it in the same pass (the bias gradient of the nn.Linear whose dy is x)."""
    _need_cuda(x, colsum_into)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    out = torch.empty((C, R), dtype=torch.bfloat16, device=x.device)
    if colsum_into is not None:
        assert colsum_into.dtype == torch.float32 and colsum_into.is_contiguous() and colsum_into.shape == (C,)
        L.check(L.load().rgqa_transpose_cast_bf16_colsum(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(out), R, C,
                                                         L.ptr(colsum_into)), "rgqa_transpose_cast_bf16_colsum")
        return out
    L.check(L.load().rgqa_transpose_cast_bf16(L.current_stream_ptr(), L.ptr(x), x.stride(0), L.ptr(out), R, C), "rgqa_transpose_cast_bf16")
    return out


def estimator_accumulate(x, target, G, class_sum, class_count):
    """One batch of the Mahalanobis estimator's data pass (src/tasks/gqa_maha.py:138-168) into fp64 device accumulators:
    x [B, D] f32, target [B, >= C] f32 multi-hot, G [D, D], class_sum [C, D], class_count [C] (all float64, zeroed by the
    caller before the first batch)."""
    _need_cuda(x, target, G, class_sum, class_count)
    assert x.dtype == torch.float32 and target.dtype == torch.float32 and x.is_contiguous() and target.stride(1) == 1
    assert G.dtype == class_sum.dtype == class_count.dtype == torch.float64
    B, D = x.shape
    C = class_count.numel()
    assert G.shape == (D, D) and class_sum.shape == (C, D) and target.shape[0] == B and target.shape[1] >= C
    assert G.is_contiguous() and class_sum.is_contiguous()
    w = torch.empty(B, dtype=torch.float32, device=x.device)
    lib = L.load()
    L.check(lib.rgqa_estimator_accumulate(L.current_stream_ptr(), L.ptr(x), L.ptr(target), target.stride(0), B, C, D, L.ptr(G),
                                          L.ptr(class_sum), L.ptr(class_count), L.ptr(w)), "rgqa_estimator_accumulate")
    return w


def score_ensemble(logits_list, method="mean"):
    """(score f32 [B], label i32 [B]) of an ensemble of checkpoints: per-class sigmoid vectors averaged ("mean") or
    multiplied ("multiply"), then arg-max — src/tasks/gqa_ensemble.py:165-177 on the device."""
    import ctypes as C
    assert method in ("mean", "multiply") and 1 <= len(logits_list) <= 8
    first = logits_list[0]
    _need_cuda(*logits_list)
    B, A = first.shape
    for t in logits_list:
        assert t.dtype == torch.float32 and t.shape == first.shape and t.stride() == first.stride() and t.stride(1) == 1
    score = torch.empty(B, dtype=torch.float32, device=first.device)
    label = torch.empty(B, dtype=torch.int32, device=first.device)
    arr = (C.c_void_p * len(logits_list))(*[L.ptr(t) for t in logits_list])
    lib = L.load()
    L.check(lib.rgqa_score_ensemble(L.current_stream_ptr(), arr, len(logits_list), 1 if method == "multiply" else 0,
                                    first.stride(0), B, A, L.ptr(score), L.ptr(label)), "rgqa_score_ensemble")
    return score, label


def score_logits(logits, mode="msp"):
    """(score f32 [B], label int32 [B]) for the logit scorers (gqa_conf.py:326, gqa_energy.py:131,185,204-206)."""
    _need_cuda(logits)
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    B, A = logits.shape
    score = torch.empty(B, dtype=torch.float32, device=logits.device)
    label = torch.empty(B, dtype=torch.int32, device=logits.device)
    lib = L.load()
    L.check(lib.rgqa_score_logits(L.current_stream_ptr(), L.SCORE_MODES[mode], L.ptr(logits), logits.stride(0), B, A,
                                  L.ptr(score), L.ptr(label)), "rgqa_score_logits")
    return score, label


def mahalanobis(x, mu, precision, return_full=False):
    """score,label = max_c -0.5 (x-mu_c)^T P (x-mu_c) (src/tasks/gqa_maha.py:228-236,270), GEMM form."""
    _need_cuda(x, mu, precision)
    B, H = x.shape
    A = mu.shape[0]
    x, mu, precision = x.contiguous(), mu.contiguous(), precision.contiguous()
    work = torch.empty(B * (H + A + 1) + 4 + A * (H + 1), dtype=torch.float32, device=x.device)
    score = torch.empty(B, dtype=torch.float32, device=x.device)
    label = torch.empty(B, dtype=torch.int32, device=x.device)
    full = torch.empty((B, A), dtype=torch.float32, device=x.device) if return_full else None
    lib = L.load()
    L.check(lib.rgqa_mahalanobis(L.current_stream_ptr(), L.ptr(x), L.ptr(mu), L.ptr(precision), B, A, H, L.ptr(work),
                                 L.ptr(score), L.ptr(label), L.ptr(full)), "rgqa_mahalanobis")
    return (score, label, full) if return_full else (score, label)


# ---- input-gradient (dgrad) kernels of the fp32 path (rgqa_b200/csrc/backward.cu) ----------------------------------
def layernorm_bwd(dy, x, gamma, add=None, eps=1e-12, scale=1.0):
    """d/dx of BertLayerNorm(x)·gamma + beta given dy (both [M, H] f32); ``add``: gradient of a parallel branch to sum in."""
    _need_cuda(dy, x, gamma, add)
    assert dy.dtype == x.dtype == torch.float32 and dy.shape == x.shape and dy.is_contiguous() and x.is_contiguous()
    assert add is None or (add.is_contiguous() and add.shape == x.shape)
    M, H = x.shape
    dx = torch.empty_like(x)
    L.check(L.load().rgqa_layernorm_bwd(L.current_stream_ptr(), L.ptr(dy), L.ptr(x), L.ptr(gamma), L.ptr(add), L.ptr(dx), M, H,
                                        eps, scale), "rgqa_layernorm_bwd")
    return dx


def layernorm_bwd_params_(dgamma, dbeta, dy, x, gamma, add=None, eps=1e-12, scale=1.0):
    """:func:`layernorm_bwd` and :func:`layernorm_param_grad_` in one pass: returns dx, ADDS to dgamma / dbeta (H = 768 / 1536)."""
    _need_cuda(dy, x, gamma, add, dgamma, dbeta)
    assert dy.dtype == x.dtype == torch.float32 and dy.shape == x.shape and dy.is_contiguous() and x.is_contiguous()
    assert add is None or (add.is_contiguous() and add.shape == x.shape)
    M, H = x.shape
    dx = torch.empty_like(x)
    L.check(L.load().rgqa_layernorm_bwd_params(L.current_stream_ptr(), L.ptr(dy), L.ptr(x), L.ptr(gamma), L.ptr(add), L.ptr(dx),
                                               L.ptr(dgamma), L.ptr(dbeta), M, H, eps, scale), "rgqa_layernorm_bwd_params")
    return dx


def gelu_fwd(x, out_dtype=torch.float32):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out_dtype == torch.bfloat16:
        y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        L.check(L.load().rgqa_gelu_fwd_bf16(L.current_stream_ptr(), L.ptr(x), L.ptr(y), x.numel()), "rgqa_gelu_fwd_bf16")
        return y
    y = torch.empty_like(x)
    L.check(L.load().rgqa_gelu_fwd(L.current_stream_ptr(), L.ptr(x), L.ptr(y), x.numel()), "rgqa_gelu_fwd")
    return y


def gelu_bwd(dy, x):
    _need_cuda(dy, x)
    assert dy.dtype == x.dtype == torch.float32 and dy.is_contiguous() and x.is_contiguous() and dy.shape == x.shape
    dx = torch.empty_like(x)
    L.check(L.load().rgqa_gelu_bwd(L.current_stream_ptr(), L.ptr(dy), L.ptr(x), L.ptr(dx), x.numel()), "rgqa_gelu_bwd")
    return dx


def tanh_bwd(dy, y):
    _need_cuda(dy, y)
    assert dy.dtype == y.dtype == torch.float32 and dy.is_contiguous() and y.is_contiguous() and dy.shape == y.shape
    dx = torch.empty_like(y)
    L.check(L.load().rgqa_tanh_bwd(L.current_stream_ptr(), L.ptr(dy), L.ptr(y), L.ptr(dx), y.numel()), "rgqa_tanh_bwd")
    return dx


def attention_dropout(qkv_q, qkv_kv, B, Lq, Lk, key_mask=None, heads=12, p=0.1, seed=0):
    """:func:`attention` (fp32) with dropout on the attention probabilities (training step); pass (p, seed) to the backward."""
    _need_cuda(qkv_q, qkv_kv, key_mask)
    H = heads * 64
    assert qkv_q.dtype == qkv_kv.dtype == torch.float32 and qkv_q.is_contiguous() and qkv_kv.is_contiguous()
    out = torch.empty((B * Lq, H), dtype=torch.float32, device=qkv_q.device)
    L.check(L.load().rgqa_attention_dropout(L.current_stream_ptr(), L.ptr(qkv_q), 3 * H, qkv_kv.data_ptr() + H * 4, 3 * H,
                                            qkv_kv.data_ptr() + 2 * H * 4, 3 * H, L.ptr(key_mask), L.ptr(out), H, B, Lq, Lk, heads,
                                            float(p), int(seed)), "rgqa_attention_dropout")
    return out


def dropout_add(x, resid=None, p=0.1, seed=0):
    """dropout(x) [+ resid] with the counter-based mask of (seed, element index); also the dropout's own backward."""
    _need_cuda(x, resid)
    assert x.dtype == torch.float32 and x.is_contiguous() and (resid is None or (resid.is_contiguous() and resid.shape == x.shape))
    y = torch.empty_like(x)
    L.check(L.load().rgqa_dropout_add(L.current_stream_ptr(), L.ptr(x), L.ptr(resid), L.ptr(y), x.numel(), float(p), int(seed)),
            "rgqa_dropout_add")
    return y


def attention_bwd(qkv_q, qkv_kv, dctx, dqkv_q, dqkv_kv, B, Lq, Lk, key_mask=None, heads=12, p=0.0, seed=0, bf16=False):
    """Backward of :func:`attention`: Q from ``qkv_q`` [B*Lq, 3H], K / V from ``qkv_kv`` [B*Lk, 3H], ``dctx`` [B*Lq, H];
    writes dQ into columns [0, H) of ``dqkv_q`` and dK, dV into columns [H, 3H) of ``dqkv_kv`` (f32, fused layout).
    ``bf16=True`` (no dropout): the five products on the tensor cores with bf16 operands (the mixed-precision training step)."""
    _need_cuda(qkv_q, qkv_kv, dctx, dqkv_q, dqkv_kv, key_mask)
    H = heads * 64
    for t in (qkv_q, qkv_kv, dctx, dqkv_q, dqkv_kv):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert qkv_q.shape == dqkv_q.shape == (B * Lq, 3 * H) and qkv_kv.shape == dqkv_kv.shape == (B * Lk, 3 * H)
    assert dctx.shape == (B * Lq, H)
    if bf16:
        assert p == 0.0, "the tensor-core attention backward has no dropout"
        L.check(L.load().rgqa_attention_bwd_bf16(L.current_stream_ptr(), L.ptr(qkv_q), 3 * H, qkv_kv.data_ptr() + H * 4, 3 * H,
                                                 qkv_kv.data_ptr() + 2 * H * 4, 3 * H, L.ptr(key_mask), L.ptr(dctx), H,
                                                 L.ptr(dqkv_q), 3 * H, dqkv_kv.data_ptr() + H * 4, 3 * H,
                                                 dqkv_kv.data_ptr() + 2 * H * 4, 3 * H, B, Lq, Lk, heads), "rgqa_attention_bwd_bf16")
        return
    L.check(L.load().rgqa_attention_bwd(L.current_stream_ptr(), L.ptr(qkv_q), 3 * H, qkv_kv.data_ptr() + H * 4, 3 * H,
                                        qkv_kv.data_ptr() + 2 * H * 4, 3 * H, L.ptr(key_mask), L.ptr(dctx), H,
                                        L.ptr(dqkv_q), 3 * H, dqkv_kv.data_ptr() + H * 4, 3 * H,
                                        dqkv_kv.data_ptr() + 2 * H * 4, 3 * H, B, Lq, Lk, heads, float(p), int(seed)), "rgqa_attention_bwd")


# ---- weight-gradient kernels (accumulating) --------------------------------------------------------------------------
def linear_wgrad_(dw, db, dy, x):
    """dw [N, K] += dy^T · x and db [N] += column sums of dy  (nn.Linear; dy [M, N], x [M, K], all f32, in place)."""
    _need_cuda(dw, db, dy, x)
    assert dy.dtype == x.dtype == dw.dtype == torch.float32 and dy.stride(1) == 1 and x.stride(1) == 1 and dw.is_contiguous()
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and dw.shape == (N, K)
    lib = L.load()
    L.check(lib.rgqa_gemm_f32_tn(L.current_stream_ptr(), L.ptr(dy), dy.stride(0), L.ptr(x), x.stride(0), L.ptr(dw), K, M, N, K),
            "rgqa_gemm_f32_tn")
    if db is not None:
        assert db.shape == (N,) and db.is_contiguous()
        L.check(lib.rgqa_colsum(L.current_stream_ptr(), L.ptr(dy), dy.stride(0), L.ptr(db), M, N), "rgqa_colsum")


def linear_wgrad_bias_(db, dy):
    """db [N] += column sums of dy [M, N] (row stride allowed)."""
    _need_cuda(db, dy)
    assert dy.dtype == db.dtype == torch.float32 and dy.stride(1) == 1 and db.is_contiguous() and db.shape == (dy.shape[1],)
    L.check(L.load().rgqa_colsum(L.current_stream_ptr(), L.ptr(dy), dy.stride(0), L.ptr(db), dy.shape[0], dy.shape[1]), "rgqa_colsum")


def layernorm_param_grad_(dgamma, dbeta, dy, x, eps=1e-12, scale=1.0):
    """dgamma += sum_m scale·dy·xh, dbeta += sum_m scale·dy for BertLayerNorm with input x [M, H] (in place)."""
    _need_cuda(dgamma, dbeta, dy, x)
    assert dy.is_contiguous() and x.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype == torch.float32
    M, H = x.shape
    L.check(L.load().rgqa_ln_param_grad(L.current_stream_ptr(), L.ptr(dy), L.ptr(x), L.ptr(dgamma), L.ptr(dbeta), M, H, eps, scale),
            "rgqa_ln_param_grad")


def embedding_scatter_(table_grad, idx, g, padding_idx=-1):
    """table_grad[idx[r], :] += g[r, :] unless idx[r] == padding_idx (nn.Embedding backward; idx int32 [rows], g [rows, H])."""
    _need_cuda(table_grad, idx, g)
    assert idx.dtype == torch.int32 and idx.is_contiguous() and g.is_contiguous() and table_grad.is_contiguous()
    rows, H = g.shape
    L.check(L.load().rgqa_embedding_scatter(L.current_stream_ptr(), L.ptr(g), L.ptr(idx), L.ptr(table_grad), rows, H,
                                            table_grad.shape[0], padding_idx), "rgqa_embedding_scatter")
