"""Fused Q|K|V + self-attention kernel, visual stream (B = 1024, S = 36, fp16 operands): seven samples per CTA pair (option
attn_straddle = 1) against three per CTA, interleaved on one box, L2 flushed between launches, medians."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
from scripts.r02_kernel_bench import head_major
lib = L.load()
dev = torch.device("cuda")
B, S, H = 1024, 36, 768
M = B * S
gen = torch.Generator().manual_seed(0)
g = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
h = (torch.randn(M, H, generator=gen) * 1.5).to(torch.float16).to(dev)
v = h.float()
st = torch.stack([v.view(M, 6, 128).sum(-1), (v * v).view(M, 6, 128).sum(-1)], -1).contiguous()
w = (torch.randn(2304, H, generator=gen) * 0.03).to(dev); bias = torch.randn(2304, generator=gen).to(dev)
wq, _, bq = ops.fold_layernorm_weight(w, bias, g, b, dtype=torch.float16)
w_hm, b_hm = head_major(wq, bq)
ctx = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
res = {0: [], 1: []}
for rep in range(14):
    for sv in (1, 0):
        lib.rgqa_set_option(b"attn_straddle", sv)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.rgqa_qkv_attention_fused_a16(L.current_stream_ptr(), h.data_ptr(), H, w_hm.data_ptr(), b_hm.data_ptr(), st.data_ptr(),
                                                 1e-12, None, ctx.data_ptr(), H, B, S, 12), "fused")
        e1.record()
        torch.cuda.synchronize()
        if rep >= 2:
            res[sv].append(e0.elapsed_time(e1))
lib.rgqa_set_option(b"attn_straddle", 0)
for sv in (0, 1):
    ms = sorted(res[sv])[len(res[sv]) // 2]
    print(f"visual Q|K|V + attention, {'7 samples / pair' if sv else '3 samples / CTA '}: {ms * 1e3:7.1f} us")
