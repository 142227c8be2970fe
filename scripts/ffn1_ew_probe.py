"""FFN1 shape (N = 3072, K = 768, bf16 output) with the erf-GELU epilogue against the plain bias epilogue and the gemm_debug
switches (bit0 no TMA stores, bit1 no epilogue math / staging), 8 vs 16 epilogue warps (option gelu_ew16): interleaved on one box,
CUDA events, L2 flushed between launches, medians.  What does the GELU cost, and where?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
lib = L.load()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
variants = [("gelu", L.EPI_GELU, 0, 0), ("gelu, 16 epilogue warps", L.EPI_GELU, 0, 1), ("gelu, 6 stages / 1 slab", L.EPI_GELU, 0, 2),
            ("bias only", L.EPI_BIAS, 0, 0), ("bias only, 6 stages / 1 slab", L.EPI_BIAS, 0, 2),
            ("gelu, no TMA stores", L.EPI_GELU, 1, 0), ("bias, no TMA stores", L.EPI_BIAS, 1, 0),
            ("no math, no staging (stores of stale smem)", L.EPI_GELU, 2, 0), ("bare main loop", L.EPI_GELU, 3, 0)]
for M in (36864, 20480):
    a = (torch.randn(M, 768, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(3072, 768, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(3072, device=dev)
    res = {v[0]: [] for v in variants}
    for rep in range(12):
        for name, epi, dbg, ew in variants:
            lib.rgqa_set_option(b"gelu_ew16", 1 if ew == 1 else 0)
            lib.rgqa_set_option(b"pair_deep", 1 if ew == 2 else 0)
            lib.rgqa_set_option(b"gemm_debug", dbg)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.linear_bf16(a, w, bias, epilogue=epi, out_dtype=torch.bfloat16)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                res[name].append(e0.elapsed_time(e1))
    for name, *_ in variants:
        ms = sorted(res[name])[len(res[name]) // 2]
        print(f"N=3072 K=768 M={M} {name:44s}: {ms * 1e3:7.1f} us  {2.0 * M * 3072 * 768 / ms / 1e9:7.1f} TFLOP/s")
lib.rgqa_set_option(b"gelu_ew16", 0)
lib.rgqa_set_option(b"pair_deep", 0)
lib.rgqa_set_option(b"gemm_debug", 0)
