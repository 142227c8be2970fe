"""The five hot kernels of the B = 1024 step, three launches each (visual-stream shapes, M = 36 864): target of ONE
`ncu --set full -k regex:"gemm_bf16_pair|attention_tma"` capture (the last launch of each kind is the warm one)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
dev = torch.device("cuda")
M, H = 36864, 768
g = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
x = torch.randn(M, H, device=dev)
hi = x.to(torch.bfloat16); lo = (x - hi.float()).to(torch.bfloat16)
st = torch.stack([x.view(M, 6, 128).sum(-1), (x * x).view(M, 6, 128).sum(-1)], -1).contiguous()
def folded(N):
    w = torch.randn(N, H, device=dev) * 0.03; bias = torch.randn(N, device=dev)
    return ops.fold_layernorm_weight(w, bias, g, b)
wq, cq, bq = folded(2304)
w1, c1, b1 = folded(3072)
wo = (torch.randn(H, H, device=dev) * 0.03).to(torch.bfloat16); bo = torch.randn(H, device=dev)
w2 = (torch.randn(H, 3072, device=dev) * 0.03).to(torch.bfloat16)
for _ in range(3):
    qkv = ops.linear_bf16_folded(hi, st, wq, bq, epilogue=L.EPI_BIAS)                       # Q|K|V consumer
for _ in range(3):
    ctx = ops.attention(qkv, qkv, 1024, 36, 36)                                            # visual self-attention
for _ in range(3):
    h2, l2, s2 = ops.linear_residual_hilo(ctx, wo, bo, hi, lo, st, g, None)                  # attention-output producer
for _ in range(3):
    inter = ops.linear_bf16_folded(h2, s2, w1, b1, epilogue=L.EPI_GELU)                    # FFN1 consumer (erf-GELU)
for _ in range(3):
    ops.linear_residual_hilo(inter, w2, bo, h2, l2, s2, g, None)                             # FFN2 producer
torch.cuda.synchronize()
print("ok")
