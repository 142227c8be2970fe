"""Where does the fused projection + attention kernel spend its time?  gemm_debug bits: 1 no attention tasks, 2 no projection
stores, 4 no context store, 8 only the first sample's tasks, 16 only the first query tile's tasks (results are wrong while set).  python scripts/r02_attn_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import _lib as L
from scripts.r02_kernel_bench import timeit, head_major

lib = L.load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B = 1024
for name, S, masked in (("visual_S36", 36, False), ("language_S20", 20, True)):
    Mr = B * S
    his = [torch.randn(Mr, 768, generator=g).to(torch.bfloat16).to(dev) for _ in range(2)]
    stats = torch.zeros(Mr, 12, device=dev); stats[:, 1] = 768.0
    w = (torch.randn(2304, 768, generator=g) * 0.03).to(torch.bfloat16).to(dev)
    bias = (torch.randn(2304, generator=g) * 0.1).to(dev)
    w_hm, b_hm = head_major(w, bias)
    lens = torch.randint(1, S + 1, (B,), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int32).to(dev) if masked else None
    ctx = torch.empty(Mr, 768, dtype=torch.bfloat16, device=dev)
    def fused(i):
        L.check(lib.rgqa_qkv_attention_fused(L.current_stream_ptr(), his[i].data_ptr(), 768, w_hm.data_ptr(), b_hm.data_ptr(),
                                             stats.data_ptr(), 1e-12, L.ptr(mask), ctx.data_ptr(), 768, B, S, 12), "fused")
    for dbg in (0, 1, 7, 8, 16, 24):
        lib.rgqa_set_option(b"gemm_debug", dbg)
        print(f"{name} debug={dbg}: {timeit(fused):8.1f} us")
    lib.rgqa_set_option(b"gemm_debug", 0)
