"""fp32 CUDA-core GEMM (precise path, backward, Mahalanobis) at a few shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops
dev = torch.device("cuda")
for M, N, K in [(1024, 768, 768), (1024, 1842, 768), (2560, 768, 768), (4608, 3072, 768), (36864, 768, 768), (20480, 2304, 768)]:
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    for _ in range(3): ops.linear_f32(a, w, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.linear_f32(a, w, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"M={M:6d} N={N:5d} K={K:4d}: {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:6.1f} TFLOP/s")
    dy = torch.randn(M, N, device=dev); dw = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
    for _ in range(2): ops.linear_wgrad_(dw, db, dy, a)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ops.linear_wgrad_(dw, db, dy, a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"   wgrad (TN + colsum)        : {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:6.1f} TFLOP/s")
