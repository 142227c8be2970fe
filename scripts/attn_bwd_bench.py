"""fp32 attention forward / backward kernels of the training and input-gradient paths at the RP-finetune shapes (B = 128 rows per
rank): CUDA-event timings, warm.  `--once` launches each backward shape three times and exits (target of an `ncu --set full -k
regex:attention_bwd` capture)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
B, H = 128, 768
once = "--once" in sys.argv
for Lq, Lk in ((36, 36), (20, 20), (20, 36), (36, 20)):
    qkv_q = torch.randn(B * Lq, 3 * H, generator=g).to(dev)
    qkv_kv = torch.randn(B * Lk, 3 * H, generator=g).to(dev) if Lk != Lq else qkv_q
    dctx = torch.randn(B * Lq, H, generator=g).to(dev)
    dq, dkv = torch.empty_like(qkv_q), torch.empty_like(qkv_kv)
    if once:
        for _ in range(3):
            ops.attention_bwd(qkv_q, qkv_kv, dctx, dq, dkv, B, Lq, Lk)
        continue
    for name, fn in (("bwd fp32 CUDA cores", lambda: ops.attention_bwd(qkv_q, qkv_kv, dctx, dq, dkv, B, Lq, Lk)),
                     ("bwd bf16 tensor cores", lambda: ops.attention_bwd(qkv_q, qkv_kv, dctx, dq, dkv, B, Lq, Lk, bf16=True)),
                     ("fwd fp32", lambda: ops.attention(qkv_q, qkv_kv, B, Lq, Lk))):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"attention {name:22s} B={B} Lq={Lq} Lk={Lk}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
torch.cuda.synchronize()
