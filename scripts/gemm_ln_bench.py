"""Fused GEMM+residual+LN kernel vs (staged RESID GEMM + LN kernel) at the LXMERT B=1024 shapes."""
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
for name, M, K in [("attn-out lang", 20480, 768), ("attn-out visn", 36864, 768), ("attn-out cross", 57344, 768),
                   ("ffn2 lang", 20480, 3072), ("ffn2 visn", 36864, 3072)]:
    N = 768
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device=dev); resid = torch.randn(M, N, device=dev)
    g = torch.ones(N, device=dev); b = torch.zeros(N, device=dev)
    t_f = timeit(lambda: ops.linear_residual_layernorm_bf16(a, w, bias, resid, g, b))
    def two():
        pre = ops.linear_bf16(a, w, bias, epilogue=L.EPI_RESID, residual=resid, out_dtype=torch.float32)
        ops.layernorm(pre, g, b, out_dtype=torch.bfloat16)
    t_2 = timeit(two)
    fl = 2.0 * M * N * K
    if os.environ.get("PROBE"):
        lib = L.load()
        for dbg in (1, 2, 4, 8, 15):
            lib.rgqa_set_option(b"gemm_debug", dbg)
            print(f"    debug={dbg:2d} (1 no stores, 2 no pass1b, 4 no resid, 8 no pass2): {timeit(lambda: ops.linear_residual_layernorm_bf16(a, w, bias, resid, g, b))*1e3:7.1f} us")
        lib.rgqa_set_option(b"gemm_debug", 0)
    print(f"{name:15s} M={M:6d} K={K:5d}: fused {t_f*1e3:7.1f} us ({fl/t_f/1e9:6.0f} TF/s) | gemm+ln {t_2*1e3:7.1f} us")
