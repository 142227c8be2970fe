"""Deferred-LayerNorm kernels vs the current (fp32-residual GEMM + LayerNorm kernel) pair at the LXMERT B=1024 shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn):
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[3]
N = 768
g = torch.ones(N, device=dev); b = torch.zeros(N, device=dev)
for name, M, K in [("attn-out lang", 20480, 768), ("attn-out visn", 36864, 768), ("ffn2 lang", 20480, 3072), ("ffn2 visn", 36864, 3072)]:
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device=dev); resid = torch.randn(M, N, device=dev)
    hi = resid.to(torch.bfloat16); lo = (resid - hi.float()).to(torch.bfloat16)
    st = torch.stack([resid.view(M, 6, 128).sum(-1), (resid * resid).view(M, 6, 128).sum(-1)], -1).contiguous()
    t_g = timeit(lambda: ops.linear_bf16(a, w, bias, epilogue=L.EPI_RESID, residual=resid, out_dtype=torch.float32))
    pre = ops.linear_bf16(a, w, bias, epilogue=L.EPI_RESID, residual=resid, out_dtype=torch.float32)
    t_l = timeit(lambda: ops.layernorm(pre, g, b, out_dtype=torch.bfloat16))
    t_h = timeit(lambda: ops.linear_residual_hilo(a, w, bias, hi, lo, st, g, b))
    fl = 2.0 * M * N * K
    print(f"producer {name:14s} M={M:6d} K={K:5d}: resid-gemm {t_g*1e3:7.1f} us + LN {t_l*1e3:6.1f} us | hilo {t_h*1e3:7.1f} us ({fl/t_h/1e9:5.0f} TF/s)")
for name, M, Nn, epi in [("qkv lang", 20480, 2304, L.EPI_BIAS), ("qkv visn", 36864, 2304, L.EPI_BIAS), ("ffn1 lang", 20480, 3072, L.EPI_GELU),
                         ("ffn1 visn", 36864, 3072, L.EPI_GELU)]:
    x = torch.randn(M, 768, device=dev); hi = x.to(torch.bfloat16)
    st = torch.stack([x.view(M, 6, 128).sum(-1), (x * x).view(M, 6, 128).sum(-1)], -1).contiguous()
    w = torch.randn(Nn, 768, device=dev) * 0.03; bias = torch.randn(Nn, device=dev)
    wf, c, b2 = ops.fold_layernorm_weight(w, bias, g, b)
    t_p = timeit(lambda: ops.linear_bf16(hi, wf, bias, epilogue=epi))
    t_f = timeit(lambda: ops.linear_bf16_folded(hi, st, wf, b2, epilogue=epi))
    fl = 2.0 * M * Nn * 768
    print(f"consumer {name:14s} M={M:6d} N={Nn:5d}: plain {t_p*1e3:7.1f} us ({fl/t_p/1e9:5.0f} TF/s) | folded {t_f*1e3:7.1f} us ({fl/t_f/1e9:5.0f} TF/s)")
