"""Time individual GEMM shapes of the LXMERT forward at B=1024 (single-CTA vs CTA-pair kernel)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
lib = L.load()
dev = torch.device("cuda")
shapes = [("qkv lang", 20480, 2304, 768, L.EPI_BIAS, torch.bfloat16), ("qkv visn", 36864, 2304, 768, L.EPI_BIAS, torch.bfloat16),
          ("ffn1 visn", 36864, 3072, 768, L.EPI_GELU, torch.bfloat16), ("attn-out visn", 36864, 768, 768, L.EPI_RESID, torch.float32),
          ("ffn2 visn", 36864, 768, 3072, L.EPI_RESID, torch.float32), ("qkv cross", 57344, 2304, 768, L.EPI_BIAS, torch.bfloat16),
          ("big-K bias", 36864, 2304, 3072, L.EPI_BIAS, torch.bfloat16)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name, M, N, K, epi, odt in shapes:
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    resid = torch.randn(M, N, device=dev) if epi == L.EPI_RESID else None
    for pair in (0, 1, 2):
        lib.rgqa_set_option(b"gemm_pair", 1 if pair else 0)
        lib.rgqa_set_option(b"gemm_staged", 1 if pair == 2 else 0)
        for _ in range(3):
            ops.linear_bf16(a, w, bias, epilogue=epi, residual=resid, out_dtype=odt)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.linear_bf16(a, w, bias, epilogue=epi, residual=resid, out_dtype=odt)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"{name:14s} M={M:6d} N={N:5d} K={K:5d} mode={['single','pair','pair+staged'][pair]:11s}: {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s")
