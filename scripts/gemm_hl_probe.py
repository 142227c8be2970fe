"""Bound analysis of the deferred-LayerNorm producer epilogue (gemm_debug bits: 1 no TMA stores, 2 no math / smem traffic,
4 no residual loads)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
lib = L.load()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
N = 768
g = torch.ones(N, device=dev); b = torch.zeros(N, device=dev)
for name, M, K in [("attn-out visn", 36864, 768), ("ffn2 visn", 36864, 3072), ("attn-out lang", 20480, 768), ("ffn2 lang", 20480, 3072)]:
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device=dev); resid = torch.randn(M, N, device=dev)
    hi = resid.to(torch.bfloat16); lo = (resid - hi.float()).to(torch.bfloat16)
    st = torch.stack([resid.view(M, 6, 128).sum(-1), (resid * resid).view(M, 6, 128).sum(-1)], -1).contiguous()
    line = f"{name:14s}"
    for dbg in (0, 1, 2, 4, 5, 7):
        lib.rgqa_set_option(b"gemm_debug", dbg)
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.linear_residual_hilo(a, w, bias, hi, lo, st, g, None); e1.record()  # the engine's variant: beta folded into the bias
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts[2:])[3]
        line += f" | dbg{dbg} {ms*1e3:6.1f}"
    print(line + " us")
lib.rgqa_set_option(b"gemm_debug", 0)
