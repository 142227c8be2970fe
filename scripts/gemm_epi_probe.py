"""Bound analysis of the staged epilogue: time a K=768 GEMM with TMA stores and/or epilogue math disabled."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
lib = L.load()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name, M, N, K, epi, odt in [("qkv visn", 36864, 2304, 768, L.EPI_BIAS, torch.bfloat16), ("ffn1 visn", 36864, 3072, 768, L.EPI_GELU, torch.bfloat16),
                               ("ffn2-shape plain bf16 out", 36864, 768, 3072, L.EPI_BIAS, torch.bfloat16), ("attnout-shape plain bf16 out", 36864, 768, 768, L.EPI_BIAS, torch.bfloat16)]:
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    for dbg in (0, 1, 2, 3):
        lib.rgqa_set_option(b"gemm_debug", dbg)
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.linear_bf16(a, w, bias, epilogue=epi, out_dtype=odt); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts[2:])[len(ts[2:]) // 2]
        print(f"{name} debug={dbg} (bit0 no TMA store, bit1 no math/smem): {ms*1e3:7.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s")
lib.rgqa_set_option(b"gemm_debug", 0)
