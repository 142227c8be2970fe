"""Time the attention kernel at B=1024 for the four (Lq, Lk) cases of the LXMERT forward."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
B, H = 1024, 768
for name, Lq, Lk in [("lang self", 20, 20), ("visn self", 36, 36), ("lang<-visn", 20, 36), ("visn<-lang", 36, 20)]:
    q = torch.randn(B * Lq, 3 * H, device=dev).to(torch.bfloat16)
    kv = q if Lq == Lk else torch.randn(B * Lk, 3 * H, device=dev).to(torch.bfloat16)
    mask = torch.ones(B, Lk, dtype=torch.int32, device=dev)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attention(q, kv, B, Lq, Lk, key_mask=mask); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[4]
    byt = (B * Lq * H * 2) * 2 + B * Lk * 2 * H * 2  # q read + out write + k,v read
    print(f"{name:11s} Lq={Lq} Lk={Lk}: {ms*1e3:7.1f} us  {byt/ms/1e6:7.1f} GB/s")
