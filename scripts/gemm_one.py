"""One GEMM shape, a few launches: target for `ncu --set full` (KIND=ffn1|qkv|attnout|ffn2|hilo_attnout|hilo_ffn2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
dev = torch.device("cuda")
kind = os.environ.get("KIND", "ffn1")
M = int(os.environ.get("M", 36864))
shape = {"ffn1": (3072, 768, L.EPI_GELU), "qkv": (2304, 768, L.EPI_BIAS), "attnout": (768, 768, L.EPI_RESID), "ffn2": (768, 3072, L.EPI_RESID),
         "hilo_attnout": (768, 768, None), "hilo_ffn2": (768, 3072, None)}[kind]
N, K, epi = shape
a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
resid = torch.randn(M, N, device=dev) if N == 768 else None
g = torch.ones(768, device=dev); b = torch.zeros(768, device=dev)
if epi is None:
    hi = resid.to(torch.bfloat16); lo = (resid - hi.float()).to(torch.bfloat16)
    st = torch.stack([resid.view(M, 6, 128).sum(-1), (resid * resid).view(M, 6, 128).sum(-1)], -1).contiguous()
for _ in range(4):
    if epi is None:
        ops.linear_residual_hilo(a, w, bias, hi, lo, st, g, b)
    elif epi == L.EPI_RESID:
        ops.linear_bf16(a, w, bias, epilogue=epi, residual=resid, out_dtype=torch.float32)
    else:
        ops.linear_bf16(a, w, bias, epilogue=epi)
torch.cuda.synchronize()
print("ok", kind, M, N, K)
