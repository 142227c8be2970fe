import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
B, H = 1024, 768
for L in (8, 16, 20, 36, 48, 64):
    q = torch.randn(B * L, 3 * H, device=dev).to(torch.bfloat16)
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attention(q, q, B, L, L); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[3]
    rd = B * L * 3 * H * 2
    print(f"L={L:3d}: {ms*1e3:7.1f} us  read {rd/ms/1e6:7.1f} GB/s  per-box {ms*1e3/ (B*12*3/148):6.3f} us")
