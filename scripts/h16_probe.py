"""fp16-state producers (attention-output K = 768, FFN2 K = 3072) at M = 36 864 / 20 480: the gemm_debug switches (bit0 no stores, bit1 no epilogue math, bit2 no residual loads), interleaved, L2 flushed, medians."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
lib = L.load()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
g = torch.ones(768, device=dev)
def run(M, K, reps=10, **opts):
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(768, K, device=dev) * 0.03).to(torch.bfloat16)
    bias = torch.randn(768, device=dev)
    h = (torch.randn(M, 768, device=dev) * 1.5).to(torch.float16)
    v = h.float()
    st = torch.stack([v.view(M, 6, 128).sum(-1), (v * v).view(M, 6, 128).sum(-1)], -1).contiguous()
    variants = [("complete", dict(gemm_debug=0)), ("no math", dict(gemm_debug=2)), ("no resid loads", dict(gemm_debug=4)),
                ("no loads, no stores", dict(gemm_debug=5)), ("bare main loop", dict(gemm_debug=7))]
    res = {n: [] for n, _ in variants}
    for rep in range(reps + 2):
        for n, o in variants:
            for k, val in o.items():
                lib.rgqa_set_option(k.encode(), val)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.linear_residual_h16(a, w, bias, h, st, g, None)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                res[n].append(e0.elapsed_time(e1))
    lib.rgqa_set_option(b"gemm_debug", 0)
    for n, _ in variants:
        ms = sorted(res[n])[len(res[n]) // 2]
        print(f"producer M={M} K={K} {n:20s}: {ms * 1e3:7.1f} us  {2.0 * M * 768 * K / ms / 1e9:7.1f} TFLOP/s")
for M in (36864, 20480):
    for K in (768, 3072):
        run(M, K)
