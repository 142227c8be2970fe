"""FFN1 (erf-GELU consumer, N = 3072, K = 768) with 8 vs 16 epilogue warps (option gelu_ew16), interleaved on one box: CUDA events,
L2 flushed between launches, medians."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
lib = L.load()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for M in (36864, 20480):
    a = (torch.randn(M, 768, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(3072, 768, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(3072, device=dev)
    outs = {}
    res = {0: [], 1: []}
    for rep in range(12):
        for ew in (0, 1):
            lib.rgqa_set_option(b"gelu_ew16", ew)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            outs[ew] = ops.linear_bf16(a, w, bias, epilogue=L.EPI_GELU, out_dtype=torch.bfloat16)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                res[ew].append(e0.elapsed_time(e1))
    assert torch.equal(outs[0], outs[1])
    for ew in (0, 1):
        ms = sorted(res[ew])[len(res[ew]) // 2]
        print(f"FFN1 M={M}: {'16' if ew else ' 8'} epilogue warps {ms * 1e3:7.1f} us  {2.0 * M * 3072 * 768 / ms / 1e9:7.1f} TFLOP/s")
lib.rgqa_set_option(b"gelu_ew16", 0)
