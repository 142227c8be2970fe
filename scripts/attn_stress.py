"""Stress the TMA attention kernel: many launches per shape, with and without mask, compared against the per-warp kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
lib = L.load()
dev = torch.device("cuda")
B, H = 1024, 768
for Lq, Lk, masked in [(36, 36, False), (36, 36, True), (20, 20, True), (20, 36, False), (36, 20, True), (64, 64, False), (48, 48, False)]:
    q = torch.randn(B * Lq, 3 * H, device=dev).to(torch.bfloat16)
    kv = q if Lq == Lk else torch.randn(B * Lk, 3 * H, device=dev).to(torch.bfloat16)
    mask = torch.ones(B, Lk, dtype=torch.int32, device=dev) if masked else None
    lib.rgqa_set_option(b"attn_tma", 0)
    ref = ops.attention(q, kv, B, Lq, Lk, key_mask=mask)
    lib.rgqa_set_option(b"attn_tma", 1)
    bad = 0
    for it in range(int(os.environ.get("ITERS", 100))):
        out = ops.attention(q, kv, B, Lq, Lk, key_mask=mask)
        if it % 10 == 0:
            bad += int(not torch.equal(out, ref))
    torch.cuda.synchronize()
    print(f"Lq={Lq} Lk={Lk} masked={masked}: ok, mismatches {bad}", flush=True)
